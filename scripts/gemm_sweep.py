import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.gemm_bench import linear_case
for (m, n, k) in [(1848, 2560, 640), (1848, 2560, 704), (1848, 2560, 768), (1848, 2560, 832), (1848, 2560, 1024),
                  (2048, 2560, 768), (8192, 2560, 768), (1848, 1280, 768), (1848, 256, 768), (128, 2560, 768), (1848, 2560, 64),
                  (1848, 2560, 128), (1848, 2560, 320)]:
    linear_case(m, n, k)
