"""Experiment: one SD-1.5 UNet forward of S samples launched kernel by kernel vs replayed from a CUDA graph (how much of the
step is inter-kernel launch gap).  python tools/graph_probe.py --samples 40"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hedit_b200

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=40)
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
cfg = dict(in_channels=4, out_channels=4, sample_size=64, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, heads=8,
           cross_attention_dim=768, norm_groups=32, ctx_len=77)
eng = hedit_b200.UNetEngine(cfg, max_samples=a.samples, max_contexts=a.samples)
eng.load_random_weights(0)
for S in sorted({8, 16, a.samples}):
    prof = eng.profile_forward(S, a.reps)
    w = eng.last_profile_whole
    print(f"S={S}: sum of kernels {sum(v[0] for v in prof.values()):.2f} ms, immediate {w['immediate']:.2f} ms, graph replay {w['graph']:.2f} ms, "
          f"{sum(v[1] for v in prof.values())} launches")
