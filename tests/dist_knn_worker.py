"""Worker for the multi-GPU kNN test: launched by torchrun (one rank per GPU, NCCL).  Every rank fits its shard of the
same synthetic bank, runs the sharded search, and rank 0 compares with the float64 oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from embeddingnet_b200 import synth  # noqa: E402
from embeddingnet_b200.models import BankKNNClassifier  # noqa: E402
from oracle import np_oracle as O  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bank, labels = synth.make_numpy(6007, 96, n_classes=70, noise=0.5)
    bank[4000] = bank[11]          # a tie across shards
    q, _ = synth.make_numpy(300, 96, seed_noise=synth.SEED_QUERY, n_classes=70, noise=0.5)
    q[0] = bank[11]
    clf = BankKNNClassifier(n_neighbors=5, process_group=dist.group.WORLD)
    clf.fit(bank, labels)          # every rank passes the full bank and keeps its slice
    d, i = clf.kneighbors(q)
    pred = clf.predict(q)
    d8, i8 = clf.kneighbors(q[:4])  # streaming path, sharded
    ok = True
    if rank == 0:
        rd, ri = O.knn_exact(bank, q, 5)
        ok = np.array_equal(i, ri) and np.allclose(d, rd, rtol=1e-5) and np.array_equal(i8, ri[:4])
        ok = ok and np.array_equal(pred, O.knn_vote(labels[ri])) and i[0, 0] == 11 and i[0, 1] == 4000
    # bank-scale semihard mining over the sharded bank: per-shard counts are all-gathered, the host draws the rank
    # (same RNG state on every rank), the owning shard resolves it
    rng = np.random.RandomState(3)
    a_idx = rng.choice(len(bank), size=40, replace=False)
    lab_ids = np.unique(labels, return_inverse=True)[1].astype(np.int32)
    pos = np.stack([bank[rng.choice(np.flatnonzero((lab_ids == lab_ids[r]) & (np.arange(len(bank)) != r)), 2,
                                    replace=False)] for r in a_idx])
    pos_d = np.sqrt(((bank[a_idx][:, None, :].astype(np.float64) - pos.astype(np.float64)) ** 2).sum(-1)
                    .astype(np.float32))
    for mode in ("semihard", "random_hard", "hardest"):
        np.random.seed(21)
        got = clf.mine_negatives(bank[a_idx], lab_ids[a_idx], positives=pos, margin=0.5, mode=mode)
        if rank == 0:
            np.random.seed(21)
            want, _ = O.mine_bank_modes(bank, lab_ids, bank[a_idx], lab_ids[a_idx], pos_d, 0.5, mode)
            ok = ok and np.array_equal(got, want)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    # all ranks must hold the identical merged result
    ids_t = torch.from_numpy(i).cuda()
    ref = ids_t.clone()
    dist.broadcast(ref, 0)
    same = bool((ref == ids_t).all().item())
    dist.barrier()
    dist.destroy_process_group()
    if not (flag.item() == 1 and same):
        print("rank %d: FAILED (oracle ok=%s, identical across ranks=%s)" % (rank, bool(flag.item()), same))
        sys.exit(1)
    if rank == 0:
        print("dist knn ok on %d ranks" % world)


if __name__ == "__main__":
    main()
