import numpy as np, time, ctypes as C, threading, sys
sys.path.insert(0,'.')
from bbtools_b200 import _lib
lib=_lib.load()
n=1_200_000_000
b=np.empty(n,np.uint8); b[:]=65
b[::7]=67; b[::5]=71
g=(n+15)//16
F=np.zeros(g,np.uint32); D=np.zeros(g,np.uint16)
def run(nt):
    per=(n//nt)//16*16
    def work(i):
        lib.bbduk_b200_pack_bases(b.ctypes.data+i*per, per, F.ctypes.data+i*per//4, D.ctypes.data+i*per//8)
    ts=[threading.Thread(target=work,args=(i,)) for i in range(nt)]
    t=time.perf_counter(); [x.start() for x in ts]; [x.join() for x in ts]; dt=time.perf_counter()-t
    return per*nt/dt/1e9
for nt in (1,2,4,8,12,16):
    run(nt); print(nt, "threads: %.1f GB/s"%max(run(nt),run(nt)))
c=np.empty_like(b)
def cp(nt):
    per=n//nt
    def work(i): np.copyto(c[i*per:(i+1)*per], b[i*per:(i+1)*per])
    ts=[threading.Thread(target=work,args=(i,)) for i in range(nt)]
    t=time.perf_counter(); [x.start() for x in ts]; [x.join() for x in ts]; dt=time.perf_counter()-t
    return per*nt/dt/1e9
for nt in (1,8,16): cp(nt); print("memcpy",nt,"threads: %.1f GB/s (read+write = 2x)"%cp(nt))
