"""Experiment: B images as two half-batches on two engines / two CUDA streams / two host threads (the tails and launch gaps of one
stream's small kernels are filled by the other's).  python tools/two_stream_bench.py --batch 8"""
import argparse, os, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hedit_b200

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--streams", type=int, default=2)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
B, T, NS = a.batch, 50, a.streams
cfg = dict(in_channels=4, out_channels=4, sample_size=64, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, heads=8,
           cross_attention_dim=768, norm_groups=32, ctx_len=77)
tok = hedit_b200.WordTokenizer()
sched = hedit_b200.DDIMTables(T, steps_offset=1)
ts, coef = hedit_b200.step_tables(sched, T, 1.0, False)
prompts = ["a green lizard is sitting on a branch", "a brown lizard is sitting on a branch"]
dev = torch.device("cuda", 0)
parts = []
Bs = B // NS
for s in range(NS):
    eng = hedit_b200.UNetEngine(cfg, max_samples=5 * Bs, max_contexts=1 + 2 * Bs)
    eng.load_random_weights(0)
    ctrls = [hedit_b200.make_controller(prompts, False, 0.4, 0.35, blend_word=(("lizard",), ("lizard",)), equilizer_params={"words": ("lizard",), "values": (2.0,)},
                                        num_steps=T, tokenizer=tok) for _ in range(Bs)]
    plan = hedit_b200.compile_edit_plan(ctrls, T)
    g = torch.Generator(device="cuda").manual_seed(s)
    parts.append(dict(eng=eng, plan=plan, xT=torch.randn(Bs, 4, 64, 64, generator=g, device=dev), zs=torch.randn(Bs, T, 4, 64, 64, generator=g, device=dev),
                      ctx=torch.randn(1 + 2 * Bs, 77, 768, generator=g, device=dev), stream=torch.cuda.Stream()))

def work(p):
    with torch.cuda.stream(p["stream"]):
        p["eng"].edit(p["xT"], p["zs"], p["ctx"], ts, coef, [1.0, 5.0, 7.5], p["plan"], 0.1, 1, False, 1)

def run_all():
    th = [threading.Thread(target=work, args=(p,)) for p in parts]
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize()

run_all()
t0 = time.perf_counter()
for _ in range(a.reps):
    run_all()
dt = (time.perf_counter() - t0) / a.reps
print(f"streams={NS} batch={B}: {B / dt:.3f} images/s ({dt * 1e3:.0f} ms per batch)")
