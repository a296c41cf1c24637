import sys; sys.path.insert(0,'.')
import numpy as np, bench
from mogpe_b200.engine import Engine
from mogpe_b200.device import DeviceBuffer
wl=bench.WORKLOADS["C4"]; spec=bench.synth_model(wl)
e,g=bench.blocks_from_spec(spec); eng=Engine(e,g,max_batch=8192)
x,y=bench.synth_batch(wl,8192,1); X=DeviceBuffer.from_numpy(x); Y=DeviceBuffer.from_numpy(y)
for i in range(2): eng.elbo(X,Y,scale=10.0,unpack=False)
eng.profile(True)
eng.elbo(X,Y,scale=10.0,unpack=False)
print(eng.profile_read())
print('Q', eng.Q, 'P', eng.P)
