import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
from oracle import cql_oracle as O
from replay_cql_b200 import layout
from replay_cql_b200.engine import CqlEngine, CqlHyperParams
from tests import helpers as Hp
for scale, squash in [(1.0,'eps'),(1e-3,'eps'),(1e-3,'softplus')]:
    B=256
    cfg=O.OracleConfig(squash=squash); st=O.init_state(cfg, seed=7)
    eng=CqlEngine(CqlHyperParams(batch_size=B, squash=squash))
    eng.set_state(Hp.oracle_state_to_flat(st))
    for step in range(3):
        batch=Hp.make_batch(B, seed=100+step, scale=scale); noise=O.make_noise(B,10,seed=200+step)
        m_ref,g_ref=O.update(cfg,st,batch,noise,True)
        m_gpu,g_gpu=eng.update_batch(Hp.batch_to_numpy(batch),Hp.noise_to_numpy(noise),True)
        print(scale,squash,step)
        for k in m_gpu: print('  ',k,m_gpu[k],m_ref[k])
        print('   g_temp',g_gpu['log_temp'],float(g_ref['log_temp']),'g_alpha',g_gpu['log_alpha'],float(g_ref['log_alpha']))
        for k in layout.NET_KEYS:
            print('   grad',k,'actor %.2e'%Hp.rel_err(g_gpu['actor'][k],g_ref['actor'][k].numpy()),' critics',['%.2e'%Hp.rel_err(g_gpu['critics'][c][k],g_ref['critics'][c][k].numpy()) for c in range(2)], 'max|g| a=%.2e c=%.2e'%(g_ref['actor'][k].abs().max(), g_ref['critics'][0][k].abs().max()))
    ref=layout.unpack_state(Hp.oracle_state_to_flat(st),2); gpu=layout.unpack_state(eng.get_state(),2)
    for k in layout.NET_KEYS:
        print('   state',k,'actor %.2e'%Hp.rel_err(gpu['actor'][k],ref['actor'][k]),'critic0 %.2e'%Hp.rel_err(gpu['critics'][0][k],ref['critics'][0][k]),'targ %.2e'%Hp.rel_err(gpu['targ_critics'][0][k],ref['targ_critics'][0][k]))
    eng.close()
# timing of sampled updates at B=1024
eng=CqlEngine(CqlHyperParams(batch_size=1024))
N=200000
rng=np.random.default_rng(0)
obs=np.stack([rng.integers(0,6040,N),rng.integers(0,3706,N)],1).astype(np.float32)
eng.load_transitions(obs, rng.integers(1,6,N).astype(np.float32), rng.integers(0,2,N).astype(np.float32), (rng.random(N)<0.01).astype(np.float32))
print(eng.update(5))
t=time.time(); m=eng.update(50); dt=time.time()-t
print('updates/s', 50/dt, m, 'launches', eng.launch_count)
