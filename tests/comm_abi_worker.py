"""Worker of test_gpu_knn.py::test_sharded_knn_through_the_c_abi_alone: one process per GPU, the exchange step of the
sharded bank kNN issued through en_comm_* (no torch.distributed).  PyTorch only provides device memory here.

    python tests/comm_abi_worker.py <rank> <nranks> <id_file> <out_file>
"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from embeddingnet_b200 import _lib, synth  # noqa: E402
from embeddingnet_b200._runtime import ptr, stream_ptr  # noqa: E402
from embeddingnet_b200.models import BankKNNClassifier  # noqa: E402


def main():
    rank, nranks, id_file, out_file = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    lib = _lib.load()
    id_buf = ctypes.create_string_buffer(_lib.EN_COMM_ID_BYTES)
    if rank == 0:
        _lib.check(lib.en_comm_unique_id(id_buf), "en_comm_unique_id")
        with open(id_file + ".tmp", "wb") as f:
            f.write(id_buf.raw)
        os.replace(id_file + ".tmp", id_file)
    else:
        t0 = time.time()
        while not os.path.exists(id_file):
            if time.time() - t0 > 120:
                raise SystemExit("rank %d: no unique id after 120 s" % rank)
            time.sleep(0.05)
        with open(id_file, "rb") as f:
            id_buf.raw = f.read()
    comm = ctypes.c_void_p()
    _lib.check(lib.en_comm_init(nranks, rank, id_buf, ctypes.byref(comm)), "en_comm_init")

    N, d, Q, k = 6000, 96, 150, 5
    bank, labels = synth.make_numpy(N, d, n_classes=60, noise=0.5)
    q, _ = synth.make_numpy(Q, d, seed_noise=synth.SEED_QUERY, n_classes=60, noise=0.5)
    lo, hi = BankKNNClassifier.shard_bounds(N, nranks, rank)
    clf = BankKNNClassifier(n_neighbors=k, device=dev).fit_shard(bank[lo:hi], labels, lo, N)   # no process group
    qd = torch.tensor(q, device=dev)
    d2, ids = clf._search(qd, k)                     # this shard's exact (d2, global id) lists
    rec = torch.stack([d2.view(torch.int64), ids]).contiguous()          # (2, Q, k) packed record
    rec_all = torch.empty((nranks, 2, Q, k), dtype=torch.int64, device=dev)
    _lib.check(lib.en_comm_allgather(comm, ptr(rec), ptr(rec_all), rec.numel() * 8, stream_ptr()), "en_comm_allgather")
    d2m = torch.empty((Q, k), dtype=torch.float64, device=dev)
    idm = torch.empty((Q, k), dtype=torch.int64, device=dev)
    _lib.call("en_knn_merge_packed", ptr(rec_all), nranks, Q, k, ptr(d2m), ptr(idm), stream_ptr())
    # mining-style exchange: every rank proposes ids for its own rows, -1 elsewhere; all-reduce(max) assembles them
    mine = torch.full((Q,), -1, dtype=torch.int64, device=dev)
    mine[rank::nranks] = torch.arange(Q, device=dev)[rank::nranks] + 1000 * rank
    got = torch.empty_like(mine)
    _lib.check(lib.en_comm_allreduce_max_i64(comm, ptr(mine), ptr(got), Q, stream_ptr()), "en_comm_allreduce_max_i64")
    torch.cuda.synchronize()
    np.savez(out_file, ids=idm.cpu().numpy(), d2=d2m.cpu().numpy(), reduced=got.cpu().numpy())
    _lib.check(lib.en_comm_destroy(comm), "en_comm_destroy")


if __name__ == "__main__":
    main()
