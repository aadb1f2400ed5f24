"""Fused optimizer step: ``torch.optim.Adam`` semantics (reference
``src/models/train.py:461``, ``adversarial_trainer.py:56-59,191``) as ONE launch over a network's flat
parameter buffer, refreshing the bf16 shadow weights in the same pass, with the global-norm clip of
``clip_grad_norm_(…, 1.0)`` (``unsupervised_trainer.py:144``) folded in as a device-side coefficient.
"""
import torch

from . import ops


class FusedAdam:
    """Adam over the flat parameter stores of uda_b200 networks (``Unet``, ``DomainDiscriminator``).

    ``step()`` expects the gradients produced by the last backward of each network (the flat buffer the
    engine filled; ``p.grad`` tensors are views of it).  ``zero_grad()`` drops them (set-to-none).
    """

    def __init__(self, networks, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=None,
                 capturable=False):
        if not isinstance(networks, (list, tuple)):
            networks = [networks]
        self.networks = list(networks)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.max_grad_norm = max_grad_norm
        self.state = {}
        self.step_count = 0
        self.last_grad_norm = None
        # capturable: the step count lives in a device int so that step() can be recorded into a CUDA graph
        self.capturable = capturable
        self._dev_step = {}

    def zero_grad(self, set_to_none=True):
        for net in self.networks:
            for p in net._store.params:
                p.grad = None
            net._store.grad = None

    def _flat_grad(self, st):
        if st.grad is None:
            raise RuntimeError("FusedAdam.step: no gradient (run backward first)")
        base = st.grad.data_ptr()
        for p in st.params:
            if p.grad is None or p.grad.data_ptr() != base + 4 * st.offsets[id(p)]:
                # autograd accumulated into a different tensor (several backward passes): gather
                g = torch.zeros_like(st.grad)
                for q in st.params:
                    if q.grad is not None:
                        st._view_like(g, q).copy_(q.grad)
                return g
        return st.grad

    @torch.no_grad()
    def step(self, grad_scale=1.0):
        self.step_count += 1
        b1, b2 = self.betas
        for net in self.networks:
            st = net._store
            g = self._flat_grad(st)
            if id(st) not in self.state:
                self.state[id(st)] = (torch.zeros_like(st.flat), torch.zeros_like(st.flat))
            m, v = self.state[id(st)]
            coef = None
            if self.max_grad_norm is not None:
                coef, self.last_grad_norm = ops.grad_clip_coef(g, self.max_grad_norm, grad_scale)
            shadow = st.shadow if getattr(net, "compute_dtype", None) == torch.bfloat16 else None
            dev_step = None
            if self.capturable:
                if id(st) not in self._dev_step:
                    self._dev_step[id(st)] = torch.full((1,), self.step_count - 1, dtype=torch.int32, device=st.flat.device)
                dev_step = self._dev_step[id(st)]
            ops.adam_step(st.flat, g, m, v, shadow, self.lr, b1, b2, self.eps, self.weight_decay, self.step_count,
                          grad_scale, coef, dev_step)
            if shadow is not None:
                st.mark_shadow_fresh()
