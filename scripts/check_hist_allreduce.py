"""K6 through the C ABI on real GPUs (run under torchrun, one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        scripts/check_hist_allreduce.py

Every rank fills a histogram with a pattern that depends on its rank, pg_hist_allreduce sums them in place over the
communicator libpagegeom.so created (pg_comm_unique_id / pg_comm_create), and the result must equal the closed form on
every rank — bit for bit, and equal to what torch.distributed's own all_reduce gives.  Rank 0 prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from multimodal_embeddings_b200 import ops
from multimodal_embeddings_b200._lib import PG_COL_HIST_BINS, PG_WIDTH_HIST_BINS, lib


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = PG_WIDTH_HIST_BINS + PG_COL_HIST_BINS
    idx = torch.arange(n, dtype=torch.int64, device="cuda")
    mine = ((idx * 2654435761 + rank * 40503) % 1000003 % 5000).to(torch.int32)
    want = sum(((idx * 2654435761 + r * 40503) % 1000003 % 5000) for r in range(world)).to(torch.int32)
    theirs = mine.clone()
    dist.all_reduce(theirs, op=dist.ReduceOp.SUM)
    hist = mine.clone()
    ops.corpus_comm()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.hist_allreduce(hist)
    b.record()
    torch.cuda.synchronize()
    ok = bool(torch.equal(hist, want)) and bool(torch.equal(hist, theirs))
    # timing of repeated exchanges (70 KB: latency-bound)
    reps = 50
    a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a2.record()
    for _ in range(reps):
        ops.hist_allreduce(hist)
    b2.record()
    torch.cuda.synchronize()
    flags = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"what": "pg_hist_allreduce over the C ABI's own NCCL communicator", "world": world, "bins": n,
                          "equal_closed_form_and_torch_all_reduce_on_every_rank": bool(flags.item()),
                          "first_call_ms": round(a.elapsed_time(b), 3), "steady_us_per_call": round(a2.elapsed_time(b2) / reps * 1e3, 1),
                          "nccl_version": int(lib().pg_comm_nccl_version())}), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
