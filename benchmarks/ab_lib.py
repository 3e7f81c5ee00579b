"""A/B harness for builds of the same source with different compile-time choices: `python benchmarks/ab_lib.py <lib.so> [bench.py args]` runs
bench.py against that build of libbpmsm.so (the package's LIB_PATH is patched in this process only; the product has no such switch)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bulletproofs_b200 as bp
bp.LIB_PATH = os.path.abspath(sys.argv[1])
sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[2:]
import bench
bench.main()
