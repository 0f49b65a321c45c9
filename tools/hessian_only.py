import sys
sys.path.insert(0, '/root/repo/tools'); sys.path.insert(0, '/root/repo')
import bench_configs as bc
bc.hessian_case(64, 50, 32, 15)
