import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_perf import sweep
sweep("C4", [int(sys.argv[1]) if len(sys.argv) > 1 else 20000], reps=3)
