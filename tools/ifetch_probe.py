import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sylow_b200
eng = sylow_b200.Engine(0)
sms = torch.cuda.get_device_properties(0).multi_processor_count
for variant, unroll in ((20, 4), (21, 16), (22, 64), (23, 256)):
    for threads, bps in ((256, 1), (128, 2), (256, 2)):
        iters = max(8, 4096 // unroll)
        best = 0
        for _ in range(3):
            ms, ops = eng.imad_probe(variant, sms * bps, threads, iters)
            best = max(best, ops / (ms * 1e-3))
        print("unroll %3d (~%4d KB body) threads=%d blocks/SM=%d: %.3f T limb-products/s" % (unroll, unroll * 3, threads, bps, best * 136 / 1e12), flush=True)
