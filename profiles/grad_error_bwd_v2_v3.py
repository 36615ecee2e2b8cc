import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from oracle import model as om
from asr_study_b200.engine import AcousticEngine, ModelSpec, pack_labels
def run(bwd):
    os.environ.pop("ASR_LSTM_BWD", None)
    if bwd: os.environ["ASR_LSTM_BWD"] = bwd
    N, T, F, H, L, C = 16, 120, 26, 512, 3, 28
    rng = np.random.RandomState(5)
    params = om.init_params(F, H, L, C, seed=9)
    x = rng.randn(N, T, F).astype(np.float32)
    lens = np.full(N, T, np.int32)
    labels = [rng.randint(0, C - 1, size=rng.randint(5, 30)).astype(np.int32) for _ in range(N)]
    eng = AcousticEngine(ModelSpec(F, H, L, C), init_params=params)
    eng.overlap = False
    flat, off, mx = pack_labels(labels, "cuda")
    feats = torch.as_tensor(np.ascontiguousarray(x.transpose(1, 0, 2))).cuda()
    eng.train_step(feats, torch.as_tensor(lens).cuda(), flat, off, mx, lr=1e-3, clipnorm=400.0)
    torch.cuda.synchronize()
    got = eng.params.export("grad")
    _, _, grads, _ = om.loss_and_grads(params, x, lens, labels, dtype=np.float64)
    errs = {k: float(np.abs(got[k] - grads[k]).max() / np.abs(grads[k]).max()) for k in grads}
    worst = max(errs, key=errs.get)
    print(bwd or "v3 (default)", "worst", worst, round(errs[worst], 5), "median", round(float(np.median(list(errs.values()))), 5),
          {k: round(v, 4) for k, v in errs.items() if k.startswith("l0.")})
run("v2"); run(None)
