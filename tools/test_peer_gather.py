"""torchrun --nproc-per-node N tools/test_peer_gather.py — the peer-memory gather of the result rows on N GPUs:
every rank feeds its own shard, waits, and checks every rank's block of the gathered rows against a local
recomputation of that rank's shard (same inputs: seeds are by rank)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import soundscope_b200 as S
from soundscope_b200.sharding import PeerGather
from bench import make_input_device

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, ch, rate, frames = 1024, 2, 48000, 19200
for mode, force in ((S.MODE_LOUDNESS, 0), (S.MODE_ALL, 0), (S.MODE_ALL, 1)):   # fused epilogue, and the separate k_results kernel
    an = S.BatchAnalyzer(n, ch, rate, mode, device=local)
    an.force_kernel(force)
    g = PeerGather(an, world=world, rank=rank, allow_p2p=os.environ.get("NO_P2P") is None)
    refs = [S.BatchAnalyzer(n, ch, rate, mode, device=local) for _ in range(world)]
    for r_ in refs:
        r_.force_kernel(force)   # same kernel as the analyzer under test: rows are compared bit for bit
    for step in range(4):
        for l in range(3):
            x = make_input_device(torch, n, frames, 100 * rank + 10 * step + l, dev)
            an.add_frames_results_device(x, g.local_rows())
            g.publish()
        rows = g.wait().clone()
        torch.cuda.synchronize()
        for r in range(world):
            for l in range(3):
                xr = make_input_device(torch, n, frames, 100 * r + 10 * step + l, dev)
                want = refs[r].add_frames_results_device(xr)
            got = rows[r * n:(r + 1) * n]
            assert torch.equal(got.nan_to_num(nan=-7.0), want.nan_to_num(nan=-7.0)), (rank, r, step, (got - want).abs().max().item())
    if rank == 0:
        print("peer gather ok:", g.kind, "mode", mode, "force", force, "epoch", an.gather_epoch() if g._p2p else "-")
    g.close()
    del an, refs
dist.barrier()
dist.destroy_process_group()
