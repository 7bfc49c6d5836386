"""Host <-> device copy rates of one node with every rank copying at once (torchrun, one rank per GPU): H2D only, D2H only, and
both directions together in cfg5's proportions (635 MB up, 197 MB down per clip).  Explains where the end-to-end numbers of
bench.py level off at 4 and 8 GPUs.   python -m torch.distributed.run --nproc-per-node N tools/probes/pcie_duplex_probe.py"""
import os
import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
UP, DOWN, REPS = 635_040_000, 197_000_000, 12
h_up = torch.empty(UP, dtype=torch.uint8).pin_memory()
h_dn = torch.empty(DOWN, dtype=torch.uint8).pin_memory()
d_up = torch.empty(UP, dtype=torch.uint8, device="cuda")
d_dn = torch.empty(DOWN, dtype=torch.uint8, device="cuda")
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()


def run(up: bool, down: bool) -> float:
    """seconds for REPS clips' worth of copies, max over ranks"""
    def once():
        if up:
            with torch.cuda.stream(s_up):
                d_up.copy_(h_up, non_blocking=True)
        if down:
            with torch.cuda.stream(s_dn):
                h_dn.copy_(d_dn, non_blocking=True)
    for _ in range(2):
        once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(REPS):
        once()
    s_up.synchronize(); s_dn.synchronize()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) * 1e-3], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


t_up, t_dn, t_both = run(True, False), run(False, True), run(True, True)
if rank == 0:
    g = lambda nbytes, t: world * REPS * nbytes / t / 1e9
    print(f"ranks {world}: H2D only {g(UP, t_up):.1f} GB/s | D2H only {g(DOWN, t_dn):.1f} GB/s | together: H2D {g(UP, t_both):.1f} + D2H {g(DOWN, t_both):.1f} "
          f"= {g(UP + DOWN, t_both):.1f} GB/s (summed over ranks; {REPS} x (635 MB up, 197 MB down) per rank)", flush=True)
if world > 1:
    dist.destroy_process_group()
