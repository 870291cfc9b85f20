import numpy as np, sys
sys.path.insert(0,'.')
import dplug_b200 as b2
from dplug_b200 import synth
for (W,H,kind,skyn) in [(620,330,'smooth',(256,192)),(620,330,'noise',(256,192)),(1920,1080,'smooth',(1024,768)),(1920,1080,'smooth',(64,48))]:
    d,m,z=synth.frame(W,H,synth.SEED,kind)
    sky=synth.skybox(*skyn)
    c=b2.PBRCompositor(0); c.resizeBuffers(W,H); c.setSkybox(sky)
    c.diffuseMap.upload(0,d); c.materialMap.upload(0,m); c.depthMap.upload(0,z)
    whole=[(0,0,W,H)]
    c.regenerateMipmaps(whole); c.compositeGUI(whole); fast=c.download()
    c.setMode(1); c.compositeGUI(whole); exact=c.download()
    diff=np.abs(fast.astype(np.int16)-exact.astype(np.int16))
    metal=m[...,1]==255
    print(W,H,kind,skyn,"max",diff.max(),"frac %.5f"%(diff>0).mean(),"frac on metal=255 px %.5f"%(diff[metal]>0).mean())
