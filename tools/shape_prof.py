"""Per-GEMM-shape times of one SD-1.5 UNet forward (HEDIT_PROFILE_SHAPES=1): ms, launches, TFLOP/s per (tag, MxNxK).
python tools/shape_prof.py --samples 40"""
import argparse, os, re, sys
os.environ["HEDIT_PROFILE_SHAPES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hedit_b200

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=40)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
cfg = dict(in_channels=4, out_channels=4, sample_size=64, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, heads=8,
           cross_attention_dim=768, norm_groups=32, ctx_len=77)
eng = hedit_b200.UNetEngine(cfg, max_samples=a.samples, max_contexts=a.samples)
eng.load_random_weights(0)
prof = eng.profile_forward(a.samples, a.reps)
tot = sum(v[0] for v in prof.values())
print(f"forward of {a.samples} samples: {tot:.2f} ms")
for tag, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0])[:45]:
    m = re.search(r"\[(\d+)x(\d+)x(\d+)\]", tag)
    tf = ""
    if m:
        M, N, K = (int(v) for v in m.groups())
        tf = f"{2.0 * M * N * K * n / ms / 1e9:7.0f} TFLOP/s"
    print(f"{ms:7.3f} ms {n:3d}x {100 * ms / tot:5.1f}%  {tag:52s} {tf}")
