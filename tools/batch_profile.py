import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
print(bench.run_batch(int(sys.argv[1]) if len(sys.argv) > 1 else 512, 512, 1024, 0, 0, 1))
