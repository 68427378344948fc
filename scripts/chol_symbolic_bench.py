"""Times the host symbolic phase of the supernodal Cholesky (nested dissection, row structures, maps; no GPU needed)\non the LF pattern of a g x g torus.  SPB_HOST_THREADS=N, SPB_ANALYZE_TIMING=2.  Usage: python scripts/chol_symbolic_bench.py 1024"""
import sys, time, ctypes as C, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sp_oracle as orc
orc.build()
import saddlepoint_b200 as spb
L = spb.capi.load()
g=int(sys.argv[1])
A = orc.lf_matrix(g, 2, 20211); p,i,x = A.csc()
n=g*g; m=2*n
stats=(C.c_int64*8)()
t=time.time()
st=L.spb_symbolic_chol_stats(m,n,p.ctypes.data_as(C.c_void_p),i.ctypes.data_as(C.c_void_p),stats,None)
print("chol_stats total", round(time.time()-t,3), st, list(stats))
