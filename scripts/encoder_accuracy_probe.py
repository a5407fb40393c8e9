import sys, numpy as np
sys.path.insert(0,'/root/repo')
from hashgan_b200.encoder import AlexNetHashEncoder, AlexNetWeights
from oracle import alexnet_oracle
w = AlexNetWeights.synthetic(64, seed=3)
img = np.random.default_rng(5).integers(0, 256, (9, 3*32*32), dtype=np.uint8)
want = alexnet_oracle.encode(img, w.tensors, 32, lrn=True)
for conv in ("fp32","tf32x3","tf32"):
    got = AlexNetHashEncoder(w, lrn=True, conv=conv)(img).cpu().numpy()
    print(conv, "max|got-oracle| =", np.abs(got-want).max(), " sign mismatches:", int(((got>0)!=(want>0)).sum()), "of", got.size, " min|h| at mismatch:", (np.abs(want)[(got>0)!=(want>0)].max() if ((got>0)!=(want>0)).any() else 0))
