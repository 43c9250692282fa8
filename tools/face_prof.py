import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hedit_b200
eng = hedit_b200.FaceUNetEngine(dict(ch=128, ch_mult=(1, 1, 2, 2, 4, 4), num_res_blocks=2, attn_resolution=16, image_size=256, in_channels=3, out_ch=3))
eng.load_random_weights(0)
x = torch.randn(8, 3, 256, 256, device="cuda")
eng(x, 500.0); torch.cuda.synchronize()
eng(x, 500.0); torch.cuda.synchronize()
