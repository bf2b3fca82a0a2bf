"""Golden vectors for the FFD transform produced by the reference's OWN data/ffd.py (build container only).

data/ffd.py imports `tensordict` and `statsmodels` (absent here).  Neither takes part in the transform
arithmetic, so this script installs two import stubs — a minimal dict-of-tensors `TensorDict` and an
`adfuller` that returns a fixed p-value — and then runs the unmodified reference class with `d_opt` given
(so `fit` evaluates `_objective` exactly once per feature, ffd.py:63-64).

    python tests/golden/make_golden_ffd.py   →  tests/golden/ffd.npz
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True


class TensorDictStub(dict):
    """The slice of tensordict.TensorDict that data/ffd.py touches (to, len, keys, [], [slice], item set)."""

    def __init__(self, data, batch_size=None):
        super().__init__(data)

    def to(self, device):
        return TensorDictStub({k: v.to(device) for k, v in self.items()})

    def __len__(self):
        return len(next(iter(self.values())))

    def __getitem__(self, key):
        if isinstance(key, slice):
            return TensorDictStub({k: v[key].clone() for k, v in self.items()})
        return dict.__getitem__(self, key)


def install_stubs():
    td = types.ModuleType("tensordict")
    td.TensorDict = TensorDictStub
    sys.modules["tensordict"] = td
    sm = types.ModuleType("statsmodels"); tsa = types.ModuleType("statsmodels.tsa")
    st = types.ModuleType("statsmodels.tsa.stattools")
    st.adfuller = lambda x, *a, **k: (0.0, 0.0495)
    sys.modules.update({"statsmodels": sm, "statsmodels.tsa": tsa, "statsmodels.tsa.stattools": st})


def main():
    install_stubs()
    sys.path.insert(0, "/root/reference")
    from data.ffd import FixedFracDiff
    torch.manual_seed(0)
    T = 3000
    rs = np.random.RandomState(5)
    names = ["open", "high", "low", "close", "ema"]
    ds = {"open": 0.4, "high": 0.25, "low": 0.6, "close": 0.9, "ema": 0.0}
    base = 100 * np.exp(np.cumsum(0.01 * rs.standard_normal((len(names), T)), axis=1))
    data = TensorDictStub({n: torch.tensor(base[i], dtype=torch.float32) for i, n in enumerate(names)})
    data["volume"] = torch.tensor(rs.random_sample(T) * 1e6, dtype=torch.float32)      # in feat_ignore → untouched
    thres = 1e-4
    ffd = FixedFracDiff(data, thres=thres, d_opt=dict(ds))
    assert str(ffd.device) == "cpu"
    out = ffd.fit_transform()
    mw = ffd.get_max_width()
    np.savez_compressed(os.path.join(HERE, "ffd.npz"),
                        names=np.array(names), d=np.array([float(ffd.d_opt[n]) for n in names]), thres=thres, T=T,
                        x=np.stack([base[i].astype(np.float32) for i in range(len(names))]),
                        widths=np.array([int(w) for w in ffd.widths], np.int32), max_width=mw,
                        out=np.stack([out[n].numpy() for n in names]),
                        volume_out=out["volume"].numpy(),
                        weights_close=ffd.weights[names.index("close")].numpy())
    print("widths", [int(w) for w in ffd.widths], "max_width", mw, "out", out["close"].shape)


if __name__ == "__main__":
    main()
