import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mb_bn
mb_bn.case(32, 16)
