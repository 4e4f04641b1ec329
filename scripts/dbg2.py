import sys; sys.path.insert(0, '.')
import numpy as np
from tests.test_gpu_descriptor_path import build_case, host_windows
from oracle import port
from nextgenmap_b200.host import CudaSW
qml, cor = 152, 27
concat, packed, reads, pairs = build_case(1234 + qml, 1500, qml, cor)
sw = CudaSW(qml, cor, lane_mode=1)
sw.set_reference(packed, len(concat)); sw.set_reads(reads)
score_buf = ((qml + cor) | 1) + 1
refs, qrys = host_windows(packed, len(concat), reads, pairs, qml, cor, score_buf, True)
for mode in (0, 1):
    want = port.batch_score(refs, qrys, qml, cor, mode); got = sw.score_pairs(mode, pairs)
    bad = np.nonzero(want != got)[0]
    print('mode', mode, 'bad', len(bad), 'concat', len(concat))
    for i in bad[:8]:
        p = pairs[i]
        print(i, 'start', int(p['window_start']), 'odd' if int(p['window_start']) & 1 else 'even', 'flags', int(p['flags']), 'read', int(p['read_index']), 'kind', int(p['read_index']) % 10, 'got', got[i], 'want', want[i])
        print('   win', refs[i].tobytes()[:60], '...', refs[i].tobytes()[-40:])
        print('   qry', qrys[i].tobytes()[:60])
