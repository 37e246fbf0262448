"""Consumers of the precompute output for BASELINE config 5 (measurement only -- the models are out of scope of this
build, SURVEY 8a "scope").  The two networks of the reference (src/model.py:5-91 CNN8, :94-202 VGG) are rebuilt here
from a layer table so that `bench.py` can time `precompute -> forward` on the GPU box, where /root/reference does not
exist.  Random-init weights, eval mode; parameter counts match the reference's (CNN8 ~2.43M, VGG ~8.15M with 39
scalars; the precompute path emits 36)."""
import torch
import torch.nn as nn


def _conv_stack(spec, cin, act, bn_first, bias):
    layers = []
    for item in spec:
        if item == "P":
            layers.append(nn.MaxPool2d(2))
        elif item == "Pc":
            layers.append(nn.MaxPool2d(2, 2, ceil_mode=True))
        elif isinstance(item, tuple):                       # (channels, stride)
            layers.append(nn.Conv2d(cin, item[0], 3, stride=item[1], padding=1, bias=bias))
            cin = item[0]
            layers += [nn.BatchNorm2d(cin), act()] if bn_first else [act(), nn.BatchNorm2d(cin)]
        else:
            layers.append(nn.Dropout2d(item))
    return nn.Sequential(*layers), cin


def _mlp(dims, act, bn_first, bias, drop_after=(), p=0.0, last_plain=False):
    layers = []
    for i in range(len(dims) - 1):
        final = last_plain and i == len(dims) - 2
        layers.append(nn.Linear(dims[i], dims[i + 1], bias=bias or final))
        if final:
            break
        layers += [nn.BatchNorm1d(dims[i + 1]), act()] if bn_first else [act(), nn.BatchNorm1d(dims[i + 1])]
        if i in drop_after:
            layers.append(nn.Dropout(p))
    return nn.Sequential(*layers)


class CNN8(nn.Module):
    def __init__(self, in_channels=9, num_scalar_features=39, dropout_rate=0.3):
        super().__init__()
        spec = [(32, 1), (64, 1), "P", (128, 1), (128, 1), "P", dropout_rate, (256, 1), (256, 1), (256, 1), (256, 1)]
        self.cnn, c = _conv_stack(spec, in_channels, nn.ReLU, bn_first=False, bias=True)
        self.pool = nn.AdaptiveAvgPool2d((1, 1))
        self.scalar_net = _mlp([num_scalar_features, 64, 64], nn.ReLU, False, True, drop_after=(0,), p=dropout_rate)
        self.classifier = _mlp([c + 64, 256, 128, 1], nn.ReLU, False, True, drop_after=(0,), p=dropout_rate, last_plain=True)

    def forward(self, features, scalars):
        x = self.pool(self.cnn(features)).flatten(1)
        return self.classifier(torch.cat([x, self.scalar_net(scalars)], dim=1)).squeeze(1)


class VGG(nn.Module):
    def __init__(self, in_channels=9, num_scalar_features=39, dropout_rate=0.2):
        super().__init__()
        d = dropout_rate
        self.block1, c = _conv_stack([(64, 1), (64, 1), (64, 2), d * 0.5], in_channels, nn.GELU, True, False)
        self.block2, c = _conv_stack([(128, 1), (128, 1), (128, 1), "Pc", d], c, nn.GELU, True, False)
        self.block3, c3 = _conv_stack([(256, 1), (256, 1), (256, 1), "Pc", d], c, nn.GELU, True, False)
        self.block4_conv, c = _conv_stack([(512, 1), (512, 1), (512, 1), d], c3, nn.GELU, True, False)
        self.block4_residual = nn.Sequential(nn.Conv2d(c3, 512, 1, bias=False), nn.BatchNorm2d(512))
        self.global_pool = nn.AdaptiveAvgPool2d((1, 1))
        self.scalar_net = _mlp([num_scalar_features, 64, 64], nn.GELU, True, False, drop_after=(0,), p=d)
        self.classifier = _mlp([c + 64, 256, 128, 1], nn.GELU, True, False, drop_after=(0, 1), p=d, last_plain=True)

    def forward(self, features, scalars):
        x = self.block3(self.block2(self.block1(features)))
        x = self.block4_conv(x) + self.block4_residual(x)
        x = self.global_pool(x).flatten(1)
        return self.classifier(torch.cat([x, self.scalar_net(scalars)], dim=1)).squeeze(1)


def n_params(m):
    return sum(p.numel() for p in m.parameters())
