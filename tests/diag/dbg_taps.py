import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import pf_test_util as U
from oracle import weights_gen as wg
ver = sys.argv[1]; opts = dict(kv.split("=") for kv in sys.argv[2:])
m, sd = U.make_model(ver)
for k, v in opts.items(): m.set_option(k, int(v))
imgs = wg.synth_images(2, 120, 160, 0)
m.inference_batch(imgs); torch.cuda.synchronize(); print("plain ok", flush=True)
m.debug_taps(True)
m.inference_batch(imgs); torch.cuda.synchronize(); print("debug forward ok", flush=True)
t = m.read_taps(); print("taps", len(t), flush=True)
