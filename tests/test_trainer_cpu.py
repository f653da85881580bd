"""CPU: host-side logic of the CUDA-graph replay path in the trainers (batch signatures, feature flattening) -
no kernel is launched here."""
import numpy as np
import torch

from rift_b200.trainer import LightningTrainer


def _feats(bs=2, A=3, T=21, Mp=4, P=20, R=2, Pr=120):
    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt)
    return {
        "agent": {"position": z(bs, A, T, 2), "heading": z(bs, A, T), "velocity": z(bs, A, T, 2), "shape": z(bs, A, T, 2),
                  "category": z(bs, A, dt=torch.int8), "valid_mask": z(bs, A, T, dt=torch.bool)},
        "map": {"point_position": z(bs, Mp, 3, P, 2), "point_vector": z(bs, Mp, 3, P, 2), "point_orientation": z(bs, Mp, 3, P),
                "polygon_center": z(bs, Mp, 3), "polygon_type": z(bs, Mp, dt=torch.int8),
                "polygon_on_route": z(bs, Mp, dt=torch.bool), "polygon_tl_status": z(bs, Mp, dt=torch.int8),
                "polygon_has_speed_limit": z(bs, Mp, dt=torch.bool), "polygon_speed_limit": z(bs, Mp),
                "valid_mask": z(bs, Mp, P, dt=torch.bool), "extra_key_the_kernels_never_read": z(1)},
        "reference_line": {"position": z(bs, R, Pr, 2), "vector": z(bs, R, Pr, 2), "orientation": z(bs, R, Pr),
                           "valid_mask": z(bs, R, Pr, dt=torch.bool)},
        "current_state": z(bs, 7),
    }


def test_flatten_batch_names_match_packed_batch_fields():
    from rift_b200 import _lib
    feats = _feats()
    batch = {"cur_pluto_feature_torch": feats, "old_group_logits_torch": torch.zeros(2, 2, 12), "not_a_tensor": 3}
    items = LightningTrainer._flatten_batch(batch, feats)
    names = [n for n, _ in items]
    struct_fields = {f for f, _ in _lib.Batch._fields_}
    feat_names = [n[2:] for n in names if n.startswith("f.")]
    assert set(feat_names) <= struct_fields, "every flattened feature must be a rift_b200_batch pointer field"
    assert len(feat_names) == 21 and "extra_key_the_kernels_never_read" not in " ".join(names)
    assert names[-1] == "b.old_group_logits_torch" and all(torch.is_tensor(t) for _, t in items)


def test_graph_signature_distinguishes_shapes_and_dtypes():
    def key(feats, extra):
        batch = {"cur_pluto_feature_torch": feats, "x_torch": extra}
        items = LightningTrainer._flatten_batch(batch, feats)
        return tuple((n, tuple(t.shape), t.dtype) for n, t in items)

    a = key(_feats(R=2), torch.zeros(2, 2, 12))
    assert a == key(_feats(R=2), torch.ones(2, 2, 12))                 # values do not matter
    assert a != key(_feats(R=3), torch.zeros(2, 3, 12))                # padded R differs -> another graph
    assert a != key(_feats(R=2), torch.zeros(2, 2, 12, dtype=torch.float64))


def test_numpy_features_are_accepted():
    feats = _feats()
    feats["current_state"] = np.zeros((2, 7), np.float64)
    items = LightningTrainer._flatten_batch({"cur_pluto_feature_torch": feats}, feats)
    assert dict(items)["f.current_state"].dtype == torch.float64


class _StubModel:
    """Just enough of PlanningModel for LightningTrainer's host logic (no device, no library)."""
    history_steps, future_steps, radius, num_modes = 21, 80, 120, 12
    ws_generation = 0

    def __init__(self):
        self.layouts = []

    def set_trainable_layers(self, layers):
        self.layouts.append(list(layers))


def test_freeze_parameters_drops_an_optimizer_built_for_the_old_layout():
    """the arenas are re-laid out by freeze_parameters, so an existing ClipAdamW (moments sized and ordered
    for the old layout, pointers into the dead arena) must not survive it; captured graphs neither."""
    tr = LightningTrainer(_StubModel(), lr=1e-4, cl_lr_decay=0.9, weight_decay=1e-5, epochs=16, warmup_epochs=3,
                          frame_rate=10, trainable_layers=["planning_decoder.pi_head"])
    tr.optimizer, tr.scheduler = object(), object()
    tr._graphs[("k",)] = {"stale": True}
    tr._graph_seen[("k",)] = 2
    tr.freeze_parameters(["planning_decoder"])
    assert tr.optimizer is None and tr.scheduler is None
    assert tr._graphs == {} and tr._graph_seen == {}
    assert tr.model.layouts == [["planning_decoder.pi_head"], ["planning_decoder"]]


def test_train_mode_request_warns_once_about_the_parity_mode():
    import warnings
    LightningTrainer._warned_train_mode = False
    tr = LightningTrainer(_StubModel(), lr=1e-4, cl_lr_decay=0.9, weight_decay=1e-5, epochs=16, warmup_epochs=3,
                          frame_rate=10, trainable_layers=[])
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        tr.train()
        tr.train()
    assert len(w) == 1 and "deterministic parity mode" in str(w[0].message)
    assert tr.training
