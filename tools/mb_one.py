import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mb_conv
mb_conv.conv_case(8, 32, 0, 64)
mb_conv.conv_case(8, 64, 0, 64)
mb_conv.conv_case(8, 32, 0, 64, stats=False)
