"""Dev probe: why does bench.measure_migration's copy-engine comparator read ~14 GB/s for the qwen2vl7b / 16-block point
when it follows the llava7b / 4096-block point (77 GB of pools freed just before)?  Prints every repetition."""
import statistics, sys, ctypes
import torch
sys.path.insert(0, ".")
import bench

_median = statistics.median
def loud_median(ts):
    print("   reps ms:", [round(t, 4) for t in ts])
    return _median(ts)
bench.statistics.median = loud_median
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
for pool, n in (("qwen2vl7b", 16), ("llava7b", 4096), ("qwen2vl7b", 16), ("qwen2vl7b", 16), ("qwen2vl7b", 256)):
    print(pool, n)
    p = bench.measure_migration(0, 1, 0, dev, pool, n, reps=3 if n == 4096 else 5)
    print("  ->", round(p["gbs_per_pair"], 1), "memcpy", round(p["memcpy_peer_gbs"], 1), p["bytes_per_request"])
    print("  mem", torch.cuda.memory_allocated() >> 20, torch.cuda.memory_reserved() >> 20)
