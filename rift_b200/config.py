"""Model configuration and the parameter table of the Pluto trajectory policy.

The table reproduces, name for name and shape for shape, the ``state_dict()`` of the
reference ``PlanningModel`` (rift/cbv/planning/pluto/model/pluto_model.py:22-120) so that
checkpoints move in both directions (training ckpt keys are ``model.<name>``,
rift/cbv/planning/pluto/pluto.py:130-137).  ``tests/test_param_table.py`` pins it against
``tests/golden/state_dict_spec.json`` which was dumped from the reference itself.
"""
from dataclasses import dataclass, field, asdict
from typing import List, Tuple, Optional

NUM_FREQ = 64          # FourierEmbedding(_, _, 64)            pluto_model.py:55
FOURIER_IN = 2 * NUM_FREQ + 1
NAT_DEPTHS = (2, 2, 2)  # NATSequenceEncoder defaults            layers/embedding.py:9-21
NAT_KERNELS = (3, 3, 5)
NAT_HEADS = (2, 4, 8)
NAT_MLP_RATIO = 3
PE_H1, PE_H2 = 128, 256  # PointsEncoder hidden widths (hard-coded) layers/embedding.py:255-269


@dataclass
class PlutoConfig:
    radius: float = 120.0
    dim: int = 128
    state_channel: int = 6
    polygon_channel: int = 6
    history_channel: int = 9
    history_steps: int = 21
    future_steps: int = 80
    encoder_depth: int = 4
    decoder_depth: int = 4
    num_heads: int = 4
    num_modes: int = 12
    # PPO value net (rift/cbv/planning/config/ppo_pluto.yaml:42-47); None = no value net
    value_hidden: Optional[Tuple[int, ...]] = None

    @property
    def ref_points(self) -> int:      # reference-line points = int(radius)  pluto_feature.py:361-402
        return int(self.radius)

    def to_dict(self):
        d = asdict(self)
        if d["value_hidden"] is not None:
            d["value_hidden"] = list(d["value_hidden"])
        return d


def pluto_small(**kw) -> PlutoConfig:
    """Reference defaults (pluto_model.py:23-44): 4 240 589 parameters."""
    return PlutoConfig(**kw)


def pluto_medium(**kw) -> PlutoConfig:
    """Not defined by the reference; SURVEY 8(d): dim 256, 8 heads, 6+6 layers: 18 991 757 parameters."""
    base = dict(dim=256, num_heads=8, encoder_depth=6, decoder_depth=6)
    base.update(kw)
    return PlutoConfig(**base)


MODEL_ZOO = {"small": pluto_small, "medium": pluto_medium}


# --------------------------------------------------------------------------------------
# parameter table
# --------------------------------------------------------------------------------------
Spec = List[Tuple[str, Tuple[int, ...], str]]   # (name, shape, kind); kind in {"f32", "i64"}


def _linear(out: Spec, p: str, n_out: int, n_in: int, bias: bool = True):
    out.append((p + ".weight", (n_out, n_in), "f32"))
    if bias:
        out.append((p + ".bias", (n_out,), "f32"))


def _norm(out: Spec, p: str, n: int):
    out.append((p + ".weight", (n,), "f32"))
    out.append((p + ".bias", (n,), "f32"))


def _batchnorm(out: Spec, p: str, n: int):
    _norm(out, p, n)
    out.append((p + ".running_mean", (n,), "f32"))
    out.append((p + ".running_var", (n,), "f32"))
    out.append((p + ".num_batches_tracked", (), "i64"))


def _mlp_layer(out: Spec, p: str, c_in: int, hidden: int, c_out: int):
    """MLPLayer: Linear -> LayerNorm -> ReLU -> Linear (layers/mlp_layer.py:4-16)."""
    _linear(out, p + ".mlp.0", hidden, c_in)
    _norm(out, p + ".mlp.1", hidden)
    _linear(out, p + ".mlp.3", c_out, hidden)


def _fourier(out: Spec, p: str, input_dim: int, hidden: int):
    """FourierEmbedding (layers/fourier_embedding.py:22-43)."""
    out.append((p + ".freqs.weight", (input_dim, NUM_FREQ), "f32"))
    for i in range(input_dim):
        _linear(out, f"{p}.mlps.{i}.0", hidden, FOURIER_IN)
        _norm(out, f"{p}.mlps.{i}.1", hidden)
        _linear(out, f"{p}.mlps.{i}.3", hidden, hidden)
    _norm(out, p + ".to_out.0", hidden)
    _linear(out, p + ".to_out.2", hidden, hidden)


def _points_encoder(out: Spec, p: str, c_in: int, c_out: int):
    """PointsEncoder (layers/embedding.py:252-296)."""
    _linear(out, p + ".first_mlp.0", PE_H1, c_in)
    _batchnorm(out, p + ".first_mlp.1", PE_H1)
    _linear(out, p + ".first_mlp.3", PE_H2, PE_H1)
    _linear(out, p + ".second_mlp.0", PE_H2, 2 * PE_H2)
    _batchnorm(out, p + ".second_mlp.1", PE_H2)
    _linear(out, p + ".second_mlp.3", c_out, PE_H2)


def _mha(out: Spec, p: str, d: int):
    """nn.MultiheadAttention with packed in-projection."""
    out.append((p + ".in_proj_weight", (3 * d, d), "f32"))
    out.append((p + ".in_proj_bias", (3 * d,), "f32"))
    _linear(out, p + ".out_proj", d, d)


def nat_dims(cfg: PlutoConfig):
    e = cfg.dim // 4
    return [e, 2 * e, 4 * e]


def _nat_encoder(out: Spec, p: str, cfg: PlutoConfig):
    """NATSequenceEncoder (layers/embedding.py:8-87)."""
    dims = nat_dims(cfg)
    out.append((p + ".embed.proj.weight", (dims[0], cfg.history_channel, 3), "f32"))
    out.append((p + ".embed.proj.bias", (dims[0],), "f32"))
    for i, d in enumerate(dims):
        for j in range(NAT_DEPTHS[i]):
            b = f"{p}.levels.{i}.blocks.{j}"
            _norm(out, b + ".norm1", d)
            out.append((b + ".attn.rpb", (NAT_HEADS[i], 2 * NAT_KERNELS[i] - 1), "f32"))
            _linear(out, b + ".attn.qkv", 3 * d, d)
            _linear(out, b + ".attn.proj", d, d)
            _norm(out, b + ".norm2", d)
            _linear(out, b + ".mlp.fc1", NAT_MLP_RATIO * d, d)
            _linear(out, b + ".mlp.fc2", d, NAT_MLP_RATIO * d)
        if i < len(dims) - 1:
            out.append((f"{p}.levels.{i}.downsample.reduction.weight", (2 * d, d, 3), "f32"))
            _norm(out, f"{p}.levels.{i}.downsample.norm", 2 * d)
    for i, d in enumerate(dims):
        _norm(out, f"{p}.norm{i}", d)
    for i, d in enumerate(dims):
        out.append((f"{p}.lateral_convs.{i}.weight", (dims[-1], d, 3), "f32"))
        out.append((f"{p}.lateral_convs.{i}.bias", (dims[-1],), "f32"))
    out.append((p + ".fpn_conv.weight", (dims[-1], dims[-1], 3), "f32"))
    out.append((p + ".fpn_conv.bias", (dims[-1],), "f32"))


def param_spec(cfg: PlutoConfig) -> Spec:
    D = cfg.dim
    T = cfg.future_steps
    s: Spec = []
    _fourier(s, "pos_emb", 3, D)
    # AgentEncoder (modules/agent_encoder.py:8-39)
    _nat_encoder(s, "agent_encoder.history_encoder", cfg)
    s.append(("agent_encoder.ego_state_emb.pos_embed", (1, cfg.state_channel, D), "f32"))
    s.append(("agent_encoder.ego_state_emb.query", (1, 1, D), "f32"))
    for i in range(cfg.state_channel):
        _linear(s, f"agent_encoder.ego_state_emb.linears.{i}", D, 1)
    _mha(s, "agent_encoder.ego_state_emb.attn", D)
    s.append(("agent_encoder.type_emb.weight", (4, D), "f32"))
    # MapEncoder (modules/map_encoder.py:8-29)
    _points_encoder(s, "map_encoder.polygon_encoder", cfg.polygon_channel + 4, D)
    _fourier(s, "map_encoder.speed_limit_emb", 1, D)
    s.append(("map_encoder.type_emb.weight", (3, D), "f32"))
    s.append(("map_encoder.on_route_emb.weight", (2, D), "f32"))
    s.append(("map_encoder.traffic_light_emb.weight", (4, D), "f32"))
    s.append(("map_encoder.unknown_speed_emb.weight", (1, D), "f32"))
    # StaticObjectsEncoder (modules/static_objects_encoder.py:8-16)
    _fourier(s, "static_objects_encoder.obj_encoder", 2, D)
    s.append(("static_objects_encoder.type_emb.weight", (4, D), "f32"))
    # encoder blocks (layers/transformer.py:41-71)
    for i in range(cfg.encoder_depth):
        b = f"encoder_blocks.{i}"
        _norm(s, b + ".norm1", D)
        _mha(s, b + ".attn", D)
        _norm(s, b + ".norm2", D)
        _linear(s, b + ".mlp.fc1", 4 * D, D)
        _linear(s, b + ".mlp.fc2", D, 4 * D)
    _norm(s, "norm", D)
    # AgentPredictor (modules/agent_predictor.py:7-15)
    for h in ("loc_predictor", "yaw_predictor", "vel_predictor"):
        _mlp_layer(s, f"agent_predictor.{h}", D, 2 * D, 2 * T)
    # PlanningDecoder (modules/planning_decoder.py:89-133): parameters first, then children
    s.append(("planning_decoder.m_emb", (1, 1, cfg.num_modes, D), "f32"))
    s.append(("planning_decoder.m_pos", (1, cfg.num_modes, D), "f32"))
    for i in range(cfg.decoder_depth):
        b = f"planning_decoder.decoder_blocks.{i}"
        _mha(s, b + ".r2r_attn", D)
        _mha(s, b + ".m2m_attn", D)
        _mha(s, b + ".cross_attn", D)
        _linear(s, b + ".ffn.0", 4 * D, D)
        _linear(s, b + ".ffn.3", D, 4 * D)
        for n in ("norm1", "norm2", "norm3", "norm4"):
            _norm(s, f"{b}.{n}", D)
    _fourier(s, "planning_decoder.r_pos_emb", 3, D)
    _points_encoder(s, "planning_decoder.r_encoder", 6, D)
    _linear(s, "planning_decoder.q_proj", D, 2 * D)
    _linear(s, "planning_decoder.cat_x_proj", D, 2 * D)
    for h in ("loc_head", "yaw_head", "vel_head"):
        _mlp_layer(s, f"planning_decoder.{h}", D, 2 * D, 2 * T)
    _mlp_layer(s, "planning_decoder.pi_head", D, D, 1)
    # hidden_proj / ref_free_decoder (pluto_model.py:97-104)
    _linear(s, "hidden_proj.0", D, D)
    _linear(s, "hidden_proj.2", D, D)
    _mlp_layer(s, "ref_free_decoder", D, 2 * D, 4 * T)
    # PPO value net: CriticPPO (rift/gym_carla/utils/net.py:355-372,420-433)
    if cfg.value_hidden is not None:
        s.append(("value_net.state_avg", (D,), "f32"))
        s.append(("value_net.state_std", (D,), "f32"))
        s.append(("value_net.value_avg", (1,), "f32"))
        s.append(("value_net.value_std", (1,), "f32"))
        dims = [D, *cfg.value_hidden, 1]
        for i in range(len(dims) - 1):
            _linear(s, f"value_net.net.{2 * i}", dims[i + 1], dims[i])
    return s


def numel(shape) -> int:
    n = 1
    for x in shape:
        n *= x
    return n


def is_buffer(name: str) -> bool:
    """BatchNorm buffers never receive gradients.  CriticPPO's state_avg/state_std/value_avg/
    value_std are nn.Parameters (rift/gym_carla/utils/net.py:360-363) which the reference's
    freeze_parameters() switches to requires_grad=True whenever 'value_net' is trainable
    (ppo_training.yaml:26-28, ppo_trainer.py:85-96), so they are parameters here too."""
    return name.endswith((".running_mean", ".running_var", ".num_batches_tracked"))
