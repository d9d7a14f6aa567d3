import numpy as np, torch, sys
sys.path.insert(0,'.')
from oracle import gp_oracle as O
from spearmint_b200.engine import GPEIEngine
eng=GPEIEngine(dtype=torch.float32)
rs=np.random.RandomState(8); S,M,F=5,1000,1; ldm=1024
mu=rs.randn(S,F,ldm).astype(np.float32); var=np.abs(rs.randn(S,ldm)).astype(np.float32)*0.5+1e-4
var[0,:10]*=1e-4; mu[0,:,:10]+=3.0
best=rs.randn(S,F).astype(np.float32)-1.0
ei,ei_sum=eng.ei_sweep(M,S,F,torch.from_numpy(mu).cuda(),torch.from_numpy(var).cuda(),ldm,torch.from_numpy(best).cuda(),None)
got=ei.cpu().numpy()[:,:M]
ref=np.zeros((S,M))
for s in range(S):
    sd=np.sqrt(var[s,:M].astype(float))[:,None]
    ref[s]=O._ei_from_moments(best[s].astype(float)[None,:], mu[s,:,:M].astype(float).T, sd).mean(1)
bad=np.argwhere(np.abs(got-ref)>1e-9*np.abs(ref)+0)
print(len(bad))
import scipy.stats as sps
for s,j in bad[:12]:
    m=float(mu[s,0,j]); v=float(var[s,j]); b=float(best[s,0]); sdv=np.sqrt(v); u=(b-m)/sdv
    print("u=%.4f got=%.6e ref=%.6e  ucdf+pdf(np)=%.6e  pdf=%.6e ucdf=%.6e"%(u,got[s,j],ref[s,j], u*sps.norm.cdf(u)+sps.norm.pdf(u), sps.norm.pdf(u), u*sps.norm.cdf(u)))
