"""Fused optimizer step: ``torch.optim.Adam`` semantics (reference
``src/models/train.py:461``, ``adversarial_trainer.py:56-59,191``) as ONE launch over a network's flat
parameter buffer, refreshing the bf16 shadow weights in the same pass, with the global-norm clip of
``clip_grad_norm_(…, 1.0)`` (``unsupervised_trainer.py:144``) folded in as a device-side coefficient.

``FusedAdam`` is a ``torch.optim.Optimizer``: the reference trainers read ``optimizer.param_groups[0]['lr']``
(``train.py:361``, ``adversarial_trainer.py:58``), checkpoint ``optimizer.state_dict()`` (``train.py:496,678``,
``trainer_phases.py:95,202,271``) and LR schedulers write ``param_groups[i]['lr']`` — all of that works unchanged.
One param group per network; its hyper-parameters are read on every step.
"""
import torch

from . import ops


class FusedAdam(torch.optim.Optimizer):
    """Adam over the flat parameter stores of uda_b200 networks (``Unet``, ``DomainDiscriminator``).

    ``step()`` expects the gradients produced by the last backward of each network (the flat buffer the
    engine filled; ``p.grad`` tensors are views of it).  ``zero_grad()`` drops them (set-to-none).
    """

    def __init__(self, networks, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=None,
                 capturable=False):
        if not isinstance(networks, (list, tuple)):
            networks = [networks]
        self.networks = list(networks)
        for net in self.networks:
            if not hasattr(net, "_store"):
                raise TypeError("FusedAdam takes uda_b200 networks (Unet, DomainDiscriminator), not parameter lists; "
                                "use torch.optim.Adam(model.parameters()) for anything else")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        groups = [{"params": list(net._store.params)} for net in self.networks]
        super().__init__(groups, defaults)
        self.max_grad_norm = max_grad_norm
        self.step_count = 0
        self.last_grad_norm = None
        # capturable: the step count lives in a device int so that step() can be recorded into a CUDA graph
        self.capturable = capturable
        self._dev_step = {}
        self._moments = {}     # id(store) -> (exp_avg, exp_avg_sq) flat fp32 buffers

    # hyper-parameters of the first group, for callers that treated them as attributes
    lr = property(lambda self: self.param_groups[0]["lr"])
    betas = property(lambda self: self.param_groups[0]["betas"])
    eps = property(lambda self: self.param_groups[0]["eps"])
    weight_decay = property(lambda self: self.param_groups[0]["weight_decay"])

    def zero_grad(self, set_to_none=True):
        for net in self.networks:
            for p in net._store.params:
                p.grad = None
            # the flat buffer itself stays allocated: a captured step (graph.GraphedStep) refills it on replay
            net._store.grad_dropped = True

    def _flat_grad(self, st):
        if st.grad is None:
            raise RuntimeError("FusedAdam.step: no gradient (run backward first)")
        base = st.grad.data_ptr()
        if all(p.grad is None for p in st.params):
            # graph replay after a user-side zero_grad(): the captured backward refilled the flat buffer but the
            # Python-side .grad views were dropped (the replay clears `grad_dropped`)
            if getattr(st, "grad_dropped", False):
                raise RuntimeError("FusedAdam.step: no gradient (run backward first)")
            return st.grad
        for p in st.params:
            if p.grad is None or p.grad.data_ptr() != base + 4 * st.offsets[id(p)]:
                # autograd accumulated into a different tensor (several backward passes): gather
                g = torch.zeros_like(st.grad)
                for q in st.params:
                    if q.grad is not None:
                        st._view_like(g, q).copy_(q.grad)
                return g
        return st.grad

    def _state_of(self, st):
        if id(st) not in self._moments:
            self._moments[id(st)] = (torch.zeros_like(st.flat), torch.zeros_like(st.flat))
        return self._moments[id(st)]

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.step_count += 1
        for net, group in zip(self.networks, self.param_groups):
            st = net._store
            b1, b2 = group["betas"]
            g = self._flat_grad(st)
            m, v = self._state_of(st)
            coef = None
            if self.max_grad_norm is not None:
                coef, self.last_grad_norm = ops.grad_clip_coef(g, self.max_grad_norm, grad_scale)
            shadow = st.shadow if getattr(net, "compute_dtype", None) == torch.bfloat16 else None
            dev_step = None
            if self.capturable:
                if id(st) not in self._dev_step:
                    self._dev_step[id(st)] = torch.full((1,), self.step_count - 1, dtype=torch.int32, device=st.flat.device)
                dev_step = self._dev_step[id(st)]
            ops.adam_step(st.flat, g, m, v, shadow, group["lr"], b1, b2, group["eps"], group["weight_decay"],
                          self.step_count, grad_scale, coef, dev_step)
            if shadow is not None:
                st.mark_shadow_fresh()
        return loss

    # -- checkpointing (reference train.py:496,678) ------------------------------------------------------------
    def state_dict(self):
        """``{"state": {group index: {"step", "exp_avg", "exp_avg_sq"}}, "param_groups": [...]}`` — the moments are the
        flat fp32 buffers of each network (layout = ``ParamStore`` order); ``step`` is read back from the device
        counter when the optimizer is capturable."""
        state = {}
        for gi, net in enumerate(self.networks):
            st = net._store
            if id(st) in self._moments:
                m, v = self._moments[id(st)]
                step = self.step_count
                if self.capturable and id(st) in self._dev_step:
                    step = int(self._dev_step[id(st)].item())
                state[gi] = {"step": step, "exp_avg": m.clone(), "exp_avg_sq": v.clone()}
        groups = [{k: v for k, v in g.items() if k != "params"} | {"params": [gi]}
                  for gi, g in enumerate(self.param_groups)]
        return {"state": state, "param_groups": groups, "max_grad_norm": self.max_grad_norm}

    def load_state_dict(self, sd):
        if len(sd["param_groups"]) != len(self.param_groups):
            raise ValueError("FusedAdam.load_state_dict: number of networks differs from the checkpoint")
        for g, saved in zip(self.param_groups, sd["param_groups"]):
            for k, v in saved.items():
                if k != "params":
                    g[k] = v
        self.max_grad_norm = sd.get("max_grad_norm", self.max_grad_norm)
        for gi, net in enumerate(self.networks):
            s = sd["state"].get(gi, sd["state"].get(str(gi)))
            if s is None:
                continue
            st = net._store
            if st.flat is None:
                raise RuntimeError("FusedAdam.load_state_dict: move the network to its CUDA device first")
            m, v = self._state_of(st)
            if s["exp_avg"].numel() != m.numel():
                raise ValueError("FusedAdam.load_state_dict: moment buffers do not match the network's parameters")
            m.copy_(s["exp_avg"]); v.copy_(s["exp_avg_sq"])
            self.step_count = int(s["step"])
            if id(st) in self._dev_step:
                self._dev_step[id(st)].fill_(self.step_count)
