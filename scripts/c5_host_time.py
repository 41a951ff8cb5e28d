"""how long does the HOST need to enqueue one config-5 step (Python + ctypes + launches)?"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mygrad_b200 as mg
from mygrad_b200 import Tensor
from mygrad_b200.nnet import conv_nd, max_pool, relu, softmax_crossentropy
import bench
dt = np.float32; B = 2048
x_dev = mg.asdevice(np.random.default_rng(0).standard_normal((B, 3, 32, 32)).astype(dt))
labels = np.random.default_rng(1).integers(0, 10, B)
params = {k: mg.asdevice(v) for k, v in bench.c5_params(dt).items()}
def step():
    p = {k: Tensor(v, copy=False) for k, v in params.items()}
    x = Tensor(x_dev, copy=False, constant=True)
    h = max_pool(relu(conv_nd(x, p["w1"], stride=1, padding=1) + p["b1"]), (2, 2), 2)
    h = max_pool(relu(conv_nd(h, p["w2"], stride=1, padding=1) + p["b2"]), (2, 2), 2)
    loss = softmax_crossentropy(h.reshape(B, -1) @ p["wd"] + p["bd"], labels)
    loss.backward()
    return loss
for _ in range(5): step()
mg.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); mg.synchronize(); t2 = time.perf_counter()
    print(f"host enqueue {1e3*(t1-t0):.3f} ms, until done {1e3*(t2-t0):.3f} ms")
