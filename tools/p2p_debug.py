"""debug helper: simulated-rank image p2p, dumps mailboxes"""
import os, sys, struct, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gravit_b200 as g
from gravit_b200 import schedulers
from scenes import library

g.init(0)
world = 2
d = library.simple_scene(320, 200)
ranks = [schedulers.build_domain_rank(d, 0, 1, flags=1)[0] for _ in range(world)]
ref, _ = ranks[0].render(d.width, d.height)
descs = [e.p2p_export(world) for e in ranks]
for r, e in enumerate(ranks):
    e.p2p_attach(r, world, b"".join(descs))
mails = [struct.unpack_from("<4Q", ds, 256)[3] for ds in descs]

def dump(tag):
    torch.cuda.synchronize()
    for r, p in enumerate(mails):
        t = schedulers.device_tensor(p, (16,), "<u4").cpu().numpy()
        print(tag, "rank", r, "go,err,cta_done,pad,done[0..3] =", t[:8].tolist(), flush=True)

for frame in range(3):
    t0 = time.time()
    for e in ranks:
        e.image_frame_p2p(d.height, 8, 0)
    t1 = time.time()
    errs = []
    for e in ranks:
        try:
            e.sync()
        except Exception as ex:
            errs.append(str(ex))
    print("frame", frame, "enqueue %.3f s, total %.3f s" % (t1 - t0, time.time() - t0), errs, flush=True)
    dump("after frame %d" % frame)
    print("equal", np.array_equal(ranks[0].fb_read(d.width, d.height), ref), flush=True)
