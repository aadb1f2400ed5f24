"""Oracle restatement of ``DomainAdaptationModel`` — TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows ``src/models/domain_model.py:4-83``: a thin wrapper holding a segmentation model and an optional
discriminator; ``forward(x, domain_adaptation=False)`` returns the segmentation logits, or ``(logits,
discriminator(x))`` when domain adaptation is requested and a discriminator exists; ``get_features`` = the
segmentation model's ``encoder(x)`` (None without an encoder); ``train/eval/to`` are forwarded to both members and
``parameters()`` is the concatenated LIST of both members' parameters.

Pinned: ``tests/test_oracle.py::test_domain_model_restatement_matches_reference`` runs this class and the reference's
own class (imported from /root/reference when present) on the same members and compares every call.
The GPU box has no reference tree, so the GPU test wraps the uda_b200 networks in THIS restatement.
"""
import torch.nn as nn


class RefDomainAdaptationModel(nn.Module):
    def __init__(self, segmentation_model, discriminator=None):          # domain_model.py:7-17
        super().__init__()
        self.segmentation_model = segmentation_model
        self.discriminator = discriminator

    def forward(self, x, domain_adaptation=False):                        # domain_model.py:19-41
        seg_pred = self.segmentation_model(x)
        if domain_adaptation and self.discriminator is not None:
            return seg_pred, self.discriminator(x)
        return seg_pred

    def get_features(self, x):                                            # domain_model.py:43-57
        if hasattr(self.segmentation_model, "encoder"):
            return self.segmentation_model.encoder(x)
        return None

    def train(self, mode=True):                                           # domain_model.py:59-64
        self.segmentation_model.train(mode)
        if self.discriminator is not None:
            self.discriminator.train(mode)
        return self

    def eval(self):                                                       # domain_model.py:66-71
        self.segmentation_model.eval()
        if self.discriminator is not None:
            self.discriminator.eval()
        return self

    def to(self, device):                                                 # domain_model.py:73-78
        self.segmentation_model = self.segmentation_model.to(device)
        if self.discriminator is not None:
            self.discriminator = self.discriminator.to(device)
        return self

    def parameters(self):                                                 # domain_model.py:80-85
        params = list(self.segmentation_model.parameters())
        if self.discriminator is not None:
            params.extend(list(self.discriminator.parameters()))
        return params
