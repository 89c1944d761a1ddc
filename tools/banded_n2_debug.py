"""2-GPU debug driver for BandedSR (torchrun --nproc-per-node 2 tools/banded_n2_debug.py): prints every stage."""
import faulthandler, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

def say(*a):
    print(f"[{os.environ.get('RANK')} {time.time() % 1000:7.2f}]", *a, file=sys.stderr, flush=True)

faulthandler.dump_traceback_later(70, exit=True, file=sys.stderr)
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
from hitsir_b200.banded import BandedSR, LocalBandedSR
from tests.helpers import build_pair
from oracle.weights import synthetic_image
stages = sys.argv[1].split(",") if len(sys.argv) > 1 else ["eager", "local", "graph"]
model, _ = build_pair((1, 1, 1), "nearest+conv", 4, "stress", 11)
model = model.to(dev)
x = synthetic_image(1, 456, 72, seed=9).to(dev)
with torch.no_grad():
    full = model(x); torch.cuda.synchronize(); say("full ok")
    if "eager" in stages:
        b = BandedSR(model); say("constructed")
        y = b.forward(x, dst_rank=None); torch.cuda.synchronize(); say("eager ok", float((y - full).abs().max()))
    if "local" in stages:
        loc = LocalBandedSR(model, world)(x); torch.cuda.synchronize(); say("local ok", float((loc - y).abs().max()) if "eager" in stages else "")
    if "graph" in stages:
        g = BandedSR(model, graphed=True); say("graph constructed")
        g1 = g.forward(x, dst_rank=None).clone(); torch.cuda.synchronize(); say("graph 1 ok")
        g2 = g.forward(x, dst_rank=None); torch.cuda.synchronize(); say("graph 2 ok", float((g1 - g2).abs().max()), float((g1 - full).abs().max()))
if 'graph' in stages:
    g.close()
faulthandler.cancel_dump_traceback_later()
torch.cuda.synchronize()
dist.destroy_process_group()
say("done")
