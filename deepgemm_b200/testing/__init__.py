from .bench import bench_events, bench_kineto, flush_l2
from .numeric import calc_diff, count_bytes
