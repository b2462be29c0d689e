import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.gemm_bench import linear_case
linear_case(98304, 2560, 320, geglu=True)
linear_case(98304, 2560, 320)             # same MMA work, plain epilogue, 2x the output bytes
linear_case(98304, 2560, 640, geglu=True)  # 2x K
linear_case(98304, 2560, 1280, geglu=True)
linear_case(98304, 1280, 320)
linear_case(98304, 256, 320)
linear_case(98304, 512, 320)
