"""How long does weight packing take?  (SURVEY.md 8f-3 suggested caching the packed tables next to the checkpoint.)
Times the 1650 hitsir_set_param calls (device-to-device copies into the parameter arena) and hitsir_finalize_weights (pack kernels,
DynamicPosBias tables, pooled relative-position bias, operand images) of the pro configuration on cuda:0."""
import ctypes, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hitsir_b200
from hitsir_b200 import _capi

dev = torch.device("cuda:0")
torch.manual_seed(0)
m = hitsir_b200.HiT_SIR(True, True, True, **hitsir_b200.PRO_KWARGS).eval().to(dev)
lib = _capi.load()
out = {}
with torch.cuda.device(dev):
    h = m._handle(dev)
    st = torch.cuda.current_stream().cuda_stream
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for name, p in m.state_dict(keep_vars=True).items():
            t = p.detach()
            _capi.check(lib.hitsir_set_param(h, name.encode(), ctypes.c_void_p(t.data_ptr()), t.numel(), ctypes.c_void_p(st)))
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        _capi.check(lib.hitsir_finalize_weights(h, ctypes.c_void_p(st)))
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        out[f"rep{rep}"] = {"set_param_ms": round((t1 - t0) * 1e3, 2), "finalize_ms": round((t2 - t1) * 1e3, 2)}
    # what a cache would have to move instead: the packed tables are ~ this many bytes (host -> device)
    free0, _ = torch.cuda.mem_get_info()
print(json.dumps(out))
