"""CPU: the C-ABI library builds, loads and exports every symbol include/rift_b200.h declares;
host-side arena / trainable / decay logic."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from rift_b200 import _lib
from rift_b200.arena import ParamArena, trainable_names, is_decay
from rift_b200.config import MODEL_ZOO, param_spec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from rift_b200.build import build
    return build()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "rift_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rift_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    h = ctypes.CDLL(built)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(h, s), f"{s} declared in include/rift_b200.h but not exported"
    assert syms == _lib.declared_symbols(), "python binding table and header disagree"
    assert _lib.lib().rift_b200_version() == 100


def test_no_torch_types_in_abi():
    src = open(os.path.join(ROOT, "include", "rift_b200.h")).read()
    assert "at::Tensor" not in src and "#include <torch" not in src and "c10::" not in src


def test_product_path_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rift_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_trainable_and_decay_rules_match_reference_groups():
    cfg = MODEL_ZOO["small"](future_steps=40)
    g = np.load(os.path.join(ROOT, "tests", "golden", "cfg1_small.npz"))
    groups = json.loads(str(g["optim_groups"]))
    names = trainable_names(cfg, ["planning_decoder.pi_head"])
    spec = {n: s for n, s, _ in param_spec(cfg)}
    decay = sorted(n for n in names if is_decay(n, tuple(spec[n])))
    no_decay = sorted(n for n in names if not is_decay(n, tuple(spec[n])))
    assert [decay, no_decay] == groups
    with pytest.raises(ValueError, match="not found in the model"):
        trainable_names(cfg, ["planning_decoder.no_such_head"])
    full = trainable_names(cfg, ["pos_emb", "agent_encoder", "map_encoder", "static_objects_encoder",
                                 "encoder_blocks", "norm", "agent_predictor", "planning_decoder", "hidden_proj",
                                 "ref_free_decoder"])
    assert len(full) == 426 and sum(is_decay(n, tuple(spec[n])) for n in full) == 137


def test_arena_layout_and_state_dict_roundtrip():
    cfg = MODEL_ZOO["small"]()
    a = ParamArena(cfg, ["planning_decoder.pi_head"], device="cpu")
    assert 0 < a.n_decay < a.n_train < a.total
    for n, o in a.offsets.items():
        assert o % 64 == 0, n
    for n in a.decay_names:
        assert a.offsets[n] < a.n_decay
    for n in a.no_decay_names:
        assert a.n_decay <= a.offsets[n] < a.n_train
    from rift_b200.synth import synth_state_dict
    sd = synth_state_dict(cfg)
    a.load_state_dict(sd)
    back = a.state_dict()
    assert list(back) == [n for n, _, _ in param_spec(cfg)]
    for n, v in sd.items():
        assert np.array_equal(back[n].numpy(), v), n


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()
