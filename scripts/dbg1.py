import sys; sys.path.insert(0, '.')
import numpy as np
from tests import util
from nextgenmap_b200.host import CudaSW
g = util.load_golden("appendix_a")
sw = CudaSW(32, 10, lane_mode=1)
for sel in ([2], [2,2,2,2], list(range(13))):
    a = sw.BatchAlign(0, g["refs"][sel], g["qrys"][sel])
    print(sel, [(x.pBuffer1, x.pBuffer2, x.PositionOffset) for x in a][:4])
a = sw.BatchAlign(1, g["refs"][[2,3,12]], g["qrys"][[2,3,12]])
print('endfree', [(x.pBuffer1, x.pBuffer2, x.PositionOffset) for x in a])
