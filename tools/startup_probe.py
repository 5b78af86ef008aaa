import time, sys, os
t0=time.perf_counter()
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from apemost_b200 import capi
t1=time.perf_counter(); print("import+numpy %.2f s"%(t1-t0))
capi.load_library(); t2=time.perf_counter(); print("dlopen lib %.2f s"%(t2-t1))
e=capi.Engine("simplesin",1,20); t3=time.perf_counter(); print("apm_gpu_create %.2f s"%(t3-t2))
data=np.loadtxt('/root/repo/tests/golden/testlc.dat'); e.set_data(data); t4=time.perf_counter(); print("set_data %.2f s"%(t4-t3))
e.set_bounds([0,4,0,-1],[3,24,1,1]); n=20
e.set_chains(0,n,params=np.tile([1.0,15.2,0.25,0.0],(n,1)),steps=np.tile([0.01,1e-4,0.01,0.01],(n,1)),beta=np.linspace(1,0.1,n))
e.run(1,100); t5=time.perf_counter(); print("first run %.2f s"%(t5-t4))
e.run(10,100); t6=time.perf_counter(); print("second run %.3f s"%(t6-t5))
e2=capi.Engine("simplesin",1,20); t7=time.perf_counter(); print("second create %.2f s"%(t7-t6))
