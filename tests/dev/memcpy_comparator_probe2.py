"""Dev probe: run bench.py's main() with every repetition of the migration legs printed to stderr."""
import statistics, sys
sys.path.insert(0, ".")
import bench
_median = statistics.median
_orig = bench.measure_migration
def loud(rank, world, local, dev, pool_name, n_move, reps=5):
    def loud_median(ts):
        print(f"   {pool_name} {n_move} reps ms:", [round(t, 4) for t in ts], file=sys.stderr)
        return _median(ts)
    bench.statistics.median = loud_median
    try:
        return _orig(rank, world, local, dev, pool_name, n_move, reps)
    finally:
        bench.statistics.median = _median
bench.measure_migration = loud
sys.argv = ["bench.py", "--steps", "5", "--warmup", "3"]
bench.main()
