import cProfile, pstats, sys, os, io
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
from align_callbacks import make_callbacks
from reveal_b200 import reveallib_ctypes as reveallib, synth
gs = synth.genomes(2, 500000, seed=1)
idx = reveallib.index()
for k, g in enumerate(gs):
    idx.addsample("s%d" % k); idx.addsequence(g.tobytes().decode())
idx.construct()
log = []
mp, ga = make_callbacks(log, minlen=20)
pr = cProfile.Profile(); pr.enable()
idx.align(mp, ga, threads=0, minl=20, minn=2)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(18); print(s.getvalue()[:3500])

import ctypes
from reveal_b200 import _native
st=(ctypes.c_int64*2)(); sec=(ctypes.c_double*2)()
_native.lib().rv_rec_stats(idx._handle(), st, sec)
print('steps small/general', st[0], st[1], 'host s', sec[0], sec[1], 'us/step', 1e6*sec[0]/max(1,st[0]), 1e6*sec[1]/max(1,st[1]))
