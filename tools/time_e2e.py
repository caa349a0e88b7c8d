"""Where does the end-to-end batch-hard step spend its time?  (developer aid, run under gpurun)

Measures, for the C3 shape (4096 x 512 fp32 = 8.4 MB each way): the pinned-memory copy rates one way, both ways at
once, and the 3-stream pipeline of bench.py with the compute replaced by nothing / by the real step."""
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
N = 200


def wall(fn, n=N):
    fn(10)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(n)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d_in = [torch.empty_like(emb) for _ in range(3)]
g_h = [torch.empty((B, D), dtype=torch.float32).pin_memory() for _ in range(3)]


def h2d(n):
    with torch.cuda.stream(s1):
        for i in range(n):
            d_in[i % 3].copy_(emb_h, non_blocking=True)


def d2h(n):
    with torch.cuda.stream(s2):
        for i in range(n):
            g_h[i % 3].copy_(d_in[i % 3], non_blocking=True)


def both(n):
    h2d(n)
    d2h(n)


mb = B * D * 4 / 1e6
for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", both)):
    ms = wall(fn)
    print("%-5s %.4f ms per 8.4 MB copy  (%.1f GB/s per direction)" % (name, ms, mb / ms))

steppers = [BatchHardStep(B, D, margin=0.5) for _ in range(3)]
s_in, s_cmp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
DEPTH = 3
l_dev = [torch.empty(B, dtype=torch.int32, device=dev) for _ in range(DEPTH)]
l_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(DEPTH)]
ev_in = [torch.cuda.Event() for _ in range(DEPTH)]
ev_cmp = [torch.cuda.Event() for _ in range(DEPTH)]
ev_out = [torch.cuda.Event() for _ in range(DEPTH)]


def pipelined(n_steps, compute, read=True):
    for i in range(n_steps + DEPTH - 1):
        if i < n_steps:
            k = i % DEPTH
            if i >= DEPTH:
                s_in.wait_event(ev_out[k])
            with torch.cuda.stream(s_in):
                d_in[k].copy_(emb_h, non_blocking=True)
                l_dev[k].copy_(lab_h, non_blocking=True)
                ev_in[k].record()
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_in[k])
                if i >= DEPTH:
                    s_cmp.wait_event(ev_out[k])
                loss, grad = compute(k)
                ev_cmp[k].record()
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp[k])
                g_h[k].copy_(grad, non_blocking=True)
                l_host[k].copy_(loss, non_blocking=True)
                ev_out[k].record()
        j = i - (DEPTH - 1)
        if j >= 0 and read:
            kj = j % DEPTH
            ev_out[kj].synchronize()
            float(l_host[kj])


zero = torch.zeros((), device=dev)
print("pipeline, no compute, no host read : %.4f ms/step" % wall(lambda n: pipelined(n, lambda k: (zero, d_in[k]), False)))
print("pipeline, no compute, host read    : %.4f ms/step" % wall(lambda n: pipelined(n, lambda k: (zero, d_in[k]))))
print("pipeline, real step, host read     : %.4f ms/step" % wall(lambda n: pipelined(n, lambda k: steppers[k].step(d_in[k], l_dev[k]))))
t0 = time.perf_counter()
for _ in range(1000):
    steppers[0].step(d_in[0], l_dev[0])
print("host cost of one step() call       : %.4f ms" % ((time.perf_counter() - t0)))
torch.cuda.synchronize()

from embeddingnet_b200.fused import BatchHardHostPipeline  # noqa: E402

for depth in (2, 3, 4):
    pipe = BatchHardHostPipeline(B, D, margin=0.5, depth=depth)
    pg = [BatchHardHostPipeline.pinned((B, D)) for _ in range(depth)]
    pl = [BatchHardHostPipeline.pinned((1,)) for _ in range(depth)]
    base = [0]

    def native(n):
        for i in range(n + depth - 1):
            if i < n:
                pipe.submit(emb_h, lab_h, pl[i % depth], pg[i % depth])
            j = i - (depth - 1)
            if j >= 0:
                pipe.wait(base[0] + j)
                float(pl[j % depth][0])
        base[0] += n

    print("native host pipe, depth %d           : %.4f ms/step" % (depth, wall(native)))
    t0 = time.perf_counter()
    for i in range(300):
        pipe.submit(emb_h, lab_h, pl[i % depth], pg[i % depth])
    dt = time.perf_counter() - t0
    torch.cuda.synchronize()
    print("   submit() alone, back to back (includes slot waits): %.4f ms" % (dt / 300 * 1e3))
    pipe.close()
