import os,sys
sys.path.insert(0,".")
import torch
from swne_b200 import NMFHandle, synth
from swne_b200._capi import PROF_NAMES
c=synth.CONFIGS["cfg5"]; m=162500
dA=synth.make_matrix_torch(c["n"],m,c["kt"],c["scale"],c["seed"])
h=NMFHandle(0); h.set_matrix(dA)
W0,H0=synth.random_init(c["n"],m,c["k"],0)
for eng in sys.argv[1:] or ("tc","simt"):
    w=h.fit(c["k"],W0,H0,max_iter=5,rel_tol=-1.0,full_traces=-1,engine=eng)
    r0=h.fit(c["k"],w.W,w.H,max_iter=10,rel_tol=-1.0,full_traces=-1,engine=eng)
    r=h.fit(c["k"],w.W,w.H,max_iter=10,rel_tol=-1.0,full_traces=-1,profile=True,engine=eng)
    print(eng, r.options["engine"], "ms/iter %.3f"%(r0.stats["loop_ms"]/10), {n: round(v/10,3) for n,v in zip(PROF_NAMES,r.stats["prof_ms"]) if v>0}, "loss", r0.target_loss[-1], "epochs", r0.average_epochs[-1])
