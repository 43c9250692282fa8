"""One decode + backward of the SD-geometry VAE decoder (target of the ncu launch list)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hedit_b200
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
eng = hedit_b200.VaeDecoderEngine(dict(latent_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2, norm_groups=32))
eng.load_random_weights(0)
z = torch.randn(B, 4, 64, 64, device="cuda") * 5
img = eng.decode_tensor(z)
eng.backward(torch.randn_like(img))
torch.cuda.synchronize()
print("MARK")
img = eng.decode_tensor(z)
eng.backward(torch.randn_like(img))
torch.cuda.synchronize()
