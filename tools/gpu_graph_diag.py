"""Which interleaving of an eager backend breaks the CUDA-graph capture of another backend's kron()?  (inputs are generated
up front: a failed capture leaves torch's CUDA generator in capture mode)"""
import os, sys, warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laplace_b200 import B200GGN, conv_engine, models  # noqa: E402

torch.manual_seed(6)
B = 32
XS = [torch.randn(B, 3, 32, 32, device="cuda") for _ in range(8)]
YS = [torch.randint(10, (B,), device="cuda") for _ in range(8)]
MODELS = [models.make("resnet18", width=16).cuda() for _ in range(12)]


def run(tag, seq, model_e, model_g, e_kw=None):
    E = B200GGN(model_e, "classification", **(e_kw or {}))
    G = B200GGN(model_g, "classification", cuda_graph=True)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        for i, who in enumerate(seq):
            (E if who == "E" else G).kron(XS[i % 8], YS[i % 8], N=500)
        msgs = [str(x.message)[60:150] for x in w if "capture" in str(x.message)]
    ok = [e["graph"] not in (None, False) for e in G._graphs.values()]
    print(f"{tag} [{seq}]: captured={ok} {msgs[:1]}", flush=True)
    return all(ok)


def main():
    conv_engine.ELEMENTWISE_MIN_BATCH = 0
    m = iter(MODELS)
    a = next(m); run("same model, interleaved", "EGEGEG", a, a)
    a, b = next(m), next(m); run("different models, interleaved", "EGEGEG", a, b)
    a = next(m); run("same model, E without conv engine", "EGEGEG", a, a, {"conv_engine": False})
    a = next(m); run("same model, E first then G", "EEEGGG", a, a)
    a = next(m); run("same model, E only right before capture", "GGEG", a, a)
    a = next(m); run("same model, E early", "GEGG", a, a)
    a = next(m); run("same model, E unfused", "EGEGEG", a, a, {"fuse_elementwise": False})
    # can the generator be recovered after a failed capture?
    for name, fn in (("manual_seed", lambda: torch.cuda.manual_seed(1)), ("set_rng_state", lambda: torch.cuda.set_rng_state(torch.cuda.get_rng_state()))):
        try:
            fn()
            torch.randn(4, device="cuda")
            print("generator usable after", name)
            break
        except Exception as e:  # noqa: BLE001
            print("generator still broken after", name, type(e).__name__, str(e)[:80])


if __name__ == "__main__":
    main()
