import sys, os, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.cpu_oracle import Oracle
from gendr_b200.cuda import generalized_renderer as ext
o = Oracle('port')
for did in (3, 7, 11, 15):
    shape = 2.0 if did in (14, 15) else 0.0
    for sign in (-1.0, 1.0):
        for x in (0.0, 0.01, 0.1, 0.5, 1.0, 2.0, 4.0, 8.0, 12.0):
            e = o.sigmoid_forward(did, sign, x * 0.1, 0.1, shape, 0.0); g = ext.sigmoid_forward(did, sign, x * 0.1, 0.1, shape, 0.0)
            e2 = o.sigmoid_backward(did, sign, x * 0.1, 0.1, shape, 0.0); g2 = ext.sigmoid_backward(did, sign, x * 0.1, 0.1, shape, 0.0)
            flag = '' if (abs(g - e) <= 2e-6 + 2e-5 * abs(e) and abs(g2 - e2) <= 1e-5 * max(1, abs(e2))) else '   <<<<'
            print(did, sign, x, 'cdf', g, e, 'pdf', g2, e2, flag)
