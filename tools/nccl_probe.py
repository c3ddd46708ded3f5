"""all-reduce timings for the DistributedClosestPoint combine step (torchrun, N ranks)"""
import os, time, torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 5_000_000
cases = [("f64 MIN n", torch.rand(n, dtype=torch.float64, device=dev), dist.ReduceOp.MIN),
         ("i64 MIN n", torch.randint(0, 8, (n,), dtype=torch.int64, device=dev), dist.ReduceOp.MIN),
         ("i64 SUM 7n", torch.randint(0, 8, (n, 7), dtype=torch.int64, device=dev), dist.ReduceOp.SUM),
         ("f64 SUM 7n", torch.rand((n, 7), dtype=torch.float64, device=dev), dist.ReduceOp.SUM),
         ("i32 SUM 14n", torch.randint(0, 8, (n, 14), dtype=torch.int32, device=dev), dist.ReduceOp.SUM)]
for name, t, op in cases:
    for _ in range(2):
        dist.all_reduce(t, op=op)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        dist.all_reduce(t, op=op)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 5 * 1e3
    if dist.get_rank() == 0:
        print("%-12s %8.2f ms  %6.1f GB/s algbw" % (name, ms, t.numel() * t.element_size() / ms / 1e6), flush=True)
dist.destroy_process_group()
