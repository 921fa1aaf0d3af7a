import sys, torch
sys.path.insert(0, ".")
from s2anet_b200.head import S2ANetHead
from s2anet_b200 import alignconv
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda:0"
g = torch.Generator().manual_seed(5)
feats = [torch.randn(2, 256, 256 // s, 256 // s, generator=g).to(DEV) for s in (8, 16, 32, 64, 128)]
head = S2ANetHead(15).eval(); head.init_synthetic(3); head = head.to(DEV)
print("calib", head.calibrate_scores(feats, 1500))
with torch.no_grad():
    a = head.forward_levels(feats)
    ra = head.get_bboxes(feats)
    alignconv._FORCE_SIMT_F32 = True
    b = head.forward_levels(feats)
    rb = head.get_bboxes(feats)
    alignconv._FORCE_SIMT_F32 = False
for lv, (oa, ob) in enumerate(zip(a, b)):
    for k, (ta, tb) in enumerate(zip(oa, ob)):
        if torch.is_tensor(ta) and k in (2, 3, 5):
            print(lv, k, tuple(ta.shape), float((ta.float() - tb.float()).abs().max()), float(tb.float().abs().max()))
for (da, la), (db, lb) in zip(ra, rb):
    print("dets", da.shape, db.shape)
    n = min(da.shape[0], db.shape[0])
    same = (da[:n] - db[:n]).abs().max(dim=1)[0]
    bad = (same > 1e-3).nonzero().flatten()
    print("first mismatch rows", bad[:5].tolist(), "of", n)
    if bad.numel():
        i = int(bad[0])
        print(da[i].tolist(), int(la[i])); print(db[i].tolist(), int(lb[i]))
