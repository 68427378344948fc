"""Times the host symbolic phase (H pattern, SpMV layout, assembly plan; no GPU needed) on the LF pattern of a
g x g torus.  SPB_HOST_THREADS=N and SPB_ANALYZE_TIMING=1|2 control threads and the stopwatch output.
Usage: python scripts/symbolic_bench.py 1024"""
import sys, time, ctypes as C, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sp_oracle as orc
orc.build()
import saddlepoint_b200 as spb
L = spb.capi.load()
g=int(sys.argv[1])
t=time.time(); A = orc.lf_matrix(g, 2, 20211); p,i,x = A.csc(); print("gen", time.time()-t)
n=g*g; m=2*n
stats=(C.c_int64*8)()
t=time.time()
st=L.spb_symbolic_plan_check(m,n,p.ctypes.data_as(C.c_void_p),i.ctypes.data_as(C.c_void_p),stats)
print("plan_check total", time.time()-t, st, list(stats))
