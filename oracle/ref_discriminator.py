"""Oracle restatement of the image-level DomainDiscriminator — TEST INFRASTRUCTURE.

Follows ``src/models/discriminator.py:4-55``: conv4x4 s2 p1 (+bias) 3->64, LeakyReLU(0.2);
then 64->128, 128->256, 256->512 each conv4x4 s2 p1 (+bias) + BatchNorm + LeakyReLU(0.2);
global average pool; Linear(512,1); Sigmoid.  Same module tree / state_dict keys
(``features.{0,2,3,5,6,8,9}``, ``classifier.2``) and default PyTorch init.
"""
import torch.nn as nn


class RefDomainDiscriminator(nn.Module):
    def __init__(self, input_channels=3):
        super().__init__()
        layers = [nn.Conv2d(input_channels, 64, 4, 2, 1), nn.LeakyReLU(0.2, inplace=True)]
        for cin in (64, 128, 256):
            layers += [nn.Conv2d(cin, cin * 2, 4, 2, 1), nn.BatchNorm2d(cin * 2),
                       nn.LeakyReLU(0.2, inplace=True)]
        self.features = nn.Sequential(*layers)
        self.classifier = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Flatten(),
                                        nn.Linear(512, 1), nn.Sigmoid())

    def forward(self, x):
        return self.classifier(self.features(x))
