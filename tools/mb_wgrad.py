import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mb_conv
for a in [(8, 64, 0, 64), (16, 32, 0, 32), (8, 32, 32, 32), (4, 128, 0, 128), (16, 16, 16, 16), (4, 64, 64, 64)]:
    mb_conv.wgrad_case(*a)
