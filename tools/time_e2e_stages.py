"""Stage durations INSIDE the 3-stream pipeline (events on each stream around each stage), for the real step and for
a sleep kernel of the same length: which stage stretches when the three overlap?  (developer aid, gpurun)"""
import sys
import time

import torch

sys.path.insert(0, ".")
from embeddingnet_b200 import _lib, synth  # noqa: E402
from embeddingnet_b200.fused import BatchHardStep  # noqa: E402

B, D = 4096, 512
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
_lib.load()
emb, labels = synth.make_device(B, D, n_classes=512, rows_per_class=8, noise=0.5, relu=True, device=dev)
emb = torch.nn.functional.normalize(emb, dim=1).contiguous()
emb_h = emb.cpu().pin_memory()
lab_h = labels.cpu().pin_memory()
DEPTH = int(sys.argv[1]) if len(sys.argv) > 1 else 3
N = 60
s_in, s_cmp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
d_in = [torch.empty_like(emb) for _ in range(DEPTH)]
l_dev = [torch.empty(B, dtype=torch.int32, device=dev) for _ in range(DEPTH)]
g_h = [torch.empty((B, D), dtype=torch.float32).pin_memory() for _ in range(DEPTH)]
l_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(DEPTH)]
steppers = [BatchHardStep(B, D, margin=0.5) for _ in range(DEPTH)]
graphs = []
for k in range(DEPTH):
    steppers[k].step(d_in[k], l_dev[k])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        steppers[k].step(d_in[k], l_dev[k])
    graphs.append(g)
zero = torch.zeros((), device=dev)
# cycles for ~70 us at ~1.9 GHz
SLEEP = int(70e-6 * 1.9e9)


def run(compute, n=N):
    T = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    ev = [[T() for _ in range(6)] for _ in range(n)]
    out_done = [None] * DEPTH
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        k = i % DEPTH
        if out_done[k] is not None:
            out_done[k].synchronize()
        with torch.cuda.stream(s_in):
            ev[i][0].record()
            d_in[k].copy_(emb_h, non_blocking=True)
            l_dev[k].copy_(lab_h, non_blocking=True)
            ev[i][1].record()
        with torch.cuda.stream(s_cmp):
            s_cmp.wait_event(ev[i][1])
            ev[i][2].record()
            loss, grad = compute(k)
            ev[i][3].record()
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev[i][3])
            ev[i][4].record()
            g_h[k].copy_(grad, non_blocking=True)
            l_host[k].copy_(loss, non_blocking=True)
            ev[i][5].record()
        out_done[k] = ev[i][5]
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    lo, hi = n // 3, n - 2
    avg = lambda f: sum(f(i) for i in range(lo, hi)) / (hi - lo)  # noqa: E731
    h2d = avg(lambda i: ev[i][0].elapsed_time(ev[i][1]))
    cmp_ = avg(lambda i: ev[i][2].elapsed_time(ev[i][3]))
    d2h = avg(lambda i: ev[i][4].elapsed_time(ev[i][5]))
    gap_in = avg(lambda i: ev[i][1].elapsed_time(ev[i + 1][0]))     # idle time of the upload stream between steps
    wait_c = avg(lambda i: ev[i][1].elapsed_time(ev[i][2]))         # upload done -> compute starts
    wait_o = avg(lambda i: ev[i][3].elapsed_time(ev[i][4]))         # compute done -> download starts
    period = avg(lambda i: ev[i][0].elapsed_time(ev[i + 1][0]))
    print("  wall %.4f ms/step | period %.4f | h2d %.4f  compute %.4f  d2h %.4f | upload idle %.4f  in->cmp %.4f  cmp->out %.4f"
          % (wall, period, h2d, cmp_, d2h, gap_in, wait_c, wait_o))


def c_none(k):
    return zero, d_in[k]


def c_sleep(k):
    torch.cuda._sleep(SLEEP)
    return zero, d_in[k]


def c_real(k):
    return steppers[k].step(d_in[k], l_dev[k])


def c_graph(k):
    graphs[k].replay()
    return steppers[k].loss, steppers[k].grad


for name, fn in (("no compute", c_none), ("sleep 70 us", c_sleep), ("real step (4 launches)", c_real),
                 ("real step (graph)", c_graph)):
    print(name)
    run(fn, 20)
    run(fn)
