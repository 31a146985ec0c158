"""EM problems of the large PSI events of an 80 M-read job (provenance of tmp_psi/big169.npz, which tests/golden/make_psi_large_events.py selects from):
the host harness (tests/emu) builds the events from the counts written by scripts/make_depth_counts.py and the problems are dumped in decreasing
order of member count.  CPU only."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
import numpy as np
import whippet_jl_b200 as wq
from wq_inputs import graphbuild
from emu import host_emu
idx = graphbuild.index_from_structure(graphbuild.load_structure(), K=9, seed=1)
z=np.load('/tmp/evprof_counts_depth80M.npz'); node,ek,ec=z['node'],z['ek'],z['ec']
he = host_emu.HostEmu(idx); he.set_counts(node, ek, ec, {}, {}); L=he.L
nrec = L.hemu_events(he.h, host_emu._p(node.astype('f8')), C.c_uint64(len(ek)), host_emu._p(np.ascontiguousarray(ek,'<u8')), host_emu._p(ec.astype('f8')), C.c_int64(101))
sz=(C.c_uint64*4)(); L.hemu_problems_flat(he.h, sz, *([None]*9))
E_,P,A,M=[int(x) for x in sz]; print(E_,P,A,M)
pp,ni,cnt,ln,ap,app,apaths,amult,tan=np.zeros(E_+1,'<u4'),np.zeros(E_,'<u4'),np.zeros(P),np.zeros(P),np.zeros(E_+1,'<u4'),np.zeros(A+1,'<u4'),np.zeros(M,'<u4'),np.zeros(A),np.zeros(E_,'u1')
L.hemu_problems_flat(he.h, sz, *[host_emu._p(a) for a in (pp,ni,cnt,ln,ap,app,apaths,amult,tan)])
# select: the big problems (by members) first
nm=app[ap[1:]]-app[ap[:-1]]; order=np.argsort(-nm.astype(np.int64))
def subset(sel):
    o=dict(path_ptr=[0],n_inc=[],count=[],length=[],amb_ptr=[0],amb_path_ptr=[0],amb_paths=[],amb_mult=[],is_tandem=[])
    for e in sel:
        p0,p1=pp[e],pp[e+1]; a0,a1=ap[e],ap[e+1]
        o['n_inc'].append(ni[e]); o['count'].append(cnt[p0:p1]); o['length'].append(ln[p0:p1]); o['path_ptr'].append(o['path_ptr'][-1]+int(p1-p0))
        base=app[a0]
        for a in range(a0,a1): o['amb_path_ptr'].append(o['amb_path_ptr'][-1]+int(app[a+1]-app[a]))
        o['amb_paths'].append(apaths[app[a0]:app[a1]]); o['amb_mult'].append(amult[a0:a1]); o['amb_ptr'].append(o['amb_ptr'][-1]+int(a1-a0)); o['is_tandem'].append(tan[e])
    cat=lambda k,dt: np.concatenate(o[k]).astype(dt) if o[k] else np.zeros(0,dt)
    return dict(path_ptr=np.array(o['path_ptr'],'<u4'),n_inc=np.array(o['n_inc'],'<u4'),count=cat('count','<f8'),length=cat('length','<f8'),amb_ptr=np.array(o['amb_ptr'],'<u4'),
                amb_path_ptr=np.array(o['amb_path_ptr'],'<u4'),amb_paths=cat('amb_paths','<u4'),amb_mult=cat('amb_mult','<f8'),is_tandem=np.array(o['is_tandem'],'u1'))
os.makedirs(os.path.join(ROOT, 'tmp_psi'),exist_ok=True)
for name,sel in (('big24',order[:24]),('big169',order[:169]),('mid',order[169:169+600])):
    d=subset(sel); np.savez_compressed(os.path.join(ROOT, 'tmp_psi', name + '.npz'),**d)
    print(name, len(sel), 'paths', int(d['path_ptr'][-1]), 'sets', int(d['amb_ptr'][-1]), 'members', int(d['amb_path_ptr'][-1]), os.path.getsize(os.path.join(ROOT, 'tmp_psi', name + '.npz')))
