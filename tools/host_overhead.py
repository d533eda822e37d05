"""Where does the host time of one fwd+bwd step go?  (GPU box)"""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import evoke_b200
from evoke_b200 import synth, _lib
from evoke_b200.ids import DeviceIds
import bench

N, D = 16384, 768
ids_np = synth.make_study_ids(N, synth.SIZES_CFG3, seed=1234)
image = torch.tensor(synth.make_embeddings(ids_np, D, seed=1235), device="cuda", requires_grad=True)
text = torch.tensor(synth.make_embeddings(ids_np, D, seed=1236), device="cuda", requires_grad=True)
ids_dev = DeviceIds(torch.from_numpy(ids_np).cuda())

def step():
    image.grad = None; text.grad = None
    loss = evoke_b200.global_alignment(image, text, ids_dev, 0.5, precision="bf16", path="tc")
    loss.backward()

def timed(k, label):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(k): step()
    host = time.perf_counter() - t0
    b.record(); torch.cuda.synchronize()
    print(f"{label:28s} device {a.elapsed_time(b)/k:7.3f} ms/step   host enqueue {host/k*1e3:7.3f} ms/step", flush=True)

for _ in range(5): step()
timed(50, "plain")
timed(50, "plain again")
for period in (0.01, 0.05, 0.2):
    s = bench.ClockSampler(0, period); s.start(); timed(50, f"nvml sampler {period*1e3:.0f} ms"); print("   ", s.stop())
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
pr.disable(); torch.cuda.synchronize()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(28); print(st.getvalue()[:6000])
