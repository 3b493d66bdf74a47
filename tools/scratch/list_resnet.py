import sys
sys.path.insert(0, "/root/repo")
from fyusenet_b200 import hostapi
import numpy as np
net = hostapi.ResNet50(device=0, batch=1)
net.load_weights(np.zeros(net.weight_floats, np.float32))
net.setup()
for l in net.layers():
    print(l)
