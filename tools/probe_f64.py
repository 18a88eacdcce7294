import ctypes, torch, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tor10_b200 as T
lib=T._lib.load()
for mode,name in ((1,'dmma'),(0,'dfma'),(2,'mixed(dmma even warps + dfma odd warps, equal FMA per warp)')):
    for it in (2048, 8192):
        tf=ctypes.c_double(0)
        T._lib.check(lib.t10_probe_f64_peak(mode,it,ctypes.byref(tf),0))
        print(name,it,round(tf.value,2),'TF')
for w in (1, 2, 3, 4, 6, 8):
    tf = ctypes.c_double(0)
    T._lib.check(lib.t10_probe_f64_occupancy(w, 8192, ctypes.byref(tf), 0))
    print("dmma-only, %d warp(s) per scheduler:" % w, round(tf.value, 2), "TF")
