import ctypes as C, sys
sys.path.insert(0,'/root/repo')
from mcmcglm_b200 import _lib
L=_lib.load(); a=C.c_double(); b=C.c_double()
_lib.check(L.cgg_debug_coarse_error(0, C.byref(a), C.byref(b)))
print("max |softplus32 - softplus| / (1+|s|) =", a.value, "at s =", b.value, " kappa=2^-21 =", 2**-21, " ratio", 2**-21/a.value)
