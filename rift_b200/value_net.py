"""``CriticPPO`` value network on the rift_b200 operators
(rift/gym_carla/utils/net.py:355-372,420-433): (state - avg) / std -> Linear+ReLU -> Linear+ReLU -> Linear
-> * value_std + value_avg -> squeeze.  Parameters are views into the policy's flat arena (entries
``value_net.*``), so they share the optimizer apply and the gradient all-reduce with the policy head.

The reference's ``freeze_parameters`` switches every ``nn.Parameter`` of a trainable module to
``requires_grad=True`` — including state_avg / state_std / value_avg / value_std — so those four receive
gradients here as well (ppo_training.yaml:26-28, ppo_trainer.py:85-96).

Forward and backward are three tiny GEMMs each (bs x dim x 256); they run through the exported C-ABI
primitives (rift_b200_op_linear / op_gemm / op_colsum / op_act_bwd) — no torch arithmetic except the
element-wise normalisation constants (4 scalars / one vector).
"""
import torch

from . import _lib


class ValueNet:
    def __init__(self, arena, prefix="value_net"):
        self.arena, self.p = arena, prefix
        names = [n for n in arena.spec if n.startswith(prefix + ".net.") and n.endswith(".weight")]
        self.layers = sorted(int(n.split(".")[2]) for n in names)        # [0, 2, 4]
        dev = arena.params.device
        self._scratch = torch.empty(148 * 1024, dtype=torch.float32, device=dev)

    def _v(self, name):
        return self.arena.view(f"{self.p}.{name}")

    def _g(self, name):
        n = f"{self.p}.{name}"
        return self.arena.grad_view(n) if n in self.arena.trainable else None

    def forward(self, state: torch.Tensor, save=False):
        """state (bs, dim) fp32 CUDA -> value (bs,)"""
        L = _lib.lib()
        x = ((state - self._v("state_avg")) / self._v("state_std")).contiguous()
        acts = [x]
        for li, idx in enumerate(self.layers):
            w, b = self._v(f"net.{idx}.weight"), self._v(f"net.{idx}.bias")
            y = torch.empty(x.shape[0], w.shape[0], dtype=torch.float32, device=x.device)
            act = 1 if li < len(self.layers) - 1 else 0
            _lib.check(L.rift_b200_op_linear(_lib.ptr(x), x.shape[0], w.shape[1], _lib.ptr(w), _lib.ptr(b), w.shape[0], act,
                                             None, _lib.ptr(y), 1, _lib.stream_ptr()), "value_net linear")
            x = y
            acts.append(x)
        raw = x
        value = (raw * self._v("value_std") + self._v("value_avg")).squeeze(1)
        if save:
            self._saved = (state, acts, raw)
        return value

    def backward(self, dvalue: torch.Tensor):
        """d loss / d value (bs,) -> gradients of every value_net.* entry (accumulated into the arena)."""
        L = _lib.lib()
        state, acts, raw = self._saved
        bs = dvalue.shape[0]
        dv = dvalue.reshape(bs, 1).contiguous()
        g = self._g("value_std")
        if g is not None:
            g += (dv * raw).sum(0)
        g = self._g("value_avg")
        if g is not None:
            g += dv.sum(0)
        dy = (dv * self._v("value_std")).contiguous()
        for li in range(len(self.layers) - 1, -1, -1):
            idx = self.layers[li]
            w = self._v(f"net.{idx}.weight")
            x, y = acts[li], acts[li + 1]
            N, K = w.shape
            if li < len(self.layers) - 1:                                  # ReLU between hidden layers
                _lib.check(L.rift_b200_op_act_bwd(_lib.ptr(y), _lib.ptr(dy), dy.numel(), 1, _lib.stream_ptr()), "act_bwd")
            gw, gb = self._g(f"net.{idx}.weight"), self._g(f"net.{idx}.bias")
            if gw is not None:      # dW += dy^T x : A(m=out,k=row) = dy[k, m], B(n=in,k=row) = x[k, n]
                _lib.check(L.rift_b200_op_gemm(_lib.ptr(dy), 1, N, _lib.ptr(x), 1, K, _lib.ptr(gw), K, N, K, bs, 1.0, 1, None, 1,
                                               _lib.stream_ptr()), "value_net wgrad")
                _lib.check(L.rift_b200_op_colsum(_lib.ptr(dy), bs, N, _lib.ptr(gb), 1, _lib.ptr(self._scratch),
                                                 _lib.stream_ptr()), "value_net bgrad")
            dx = torch.empty(bs, K, dtype=torch.float32, device=dy.device)   # dx = dy W : B(n=in,k=out) = W[k, n]
            _lib.check(L.rift_b200_op_gemm(_lib.ptr(dy), N, 1, _lib.ptr(w), 1, K, _lib.ptr(dx), K, bs, K, N, 0.0, 1, None, 1,
                                           _lib.stream_ptr()), "value_net dgrad")
            dy = dx
        std, avg = self._v("state_std"), self._v("state_avg")
        g = self._g("state_avg")
        if g is not None:
            g += (-dy / std).sum(0)
        g = self._g("state_std")
        if g is not None:
            g += (-dy * (state - avg) / (std * std)).sum(0)
