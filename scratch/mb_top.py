import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mb_conv
mb_conv.conv_case(32, 16, 0, 16)
mb_conv.wgrad_case(32, 16, 0, 16)
mb_conv.conv_case(16, 32, 0, 32)
mb_conv.wgrad_case(16, 32, 0, 32)
