"""Drop-in for reference models/backbone.py: `resnet18(modality, args)` returns a module with the
reference's attribute tree (conv1, bn1, relu, maxpool, layer1..4 of BasicBlock(conv1, bn1, relu,
conv2, bn2, downsample)), hence identical parameter / buffer names, shapes, construction order
(same RNG stream under setup_seed) and state_dict — but whose forward runs on the sm_100a
kernels of libgdl_b200.so through an EncoderEngine instead of cuDNN.

The nn.Conv2d / nn.BatchNorm2d children are parameter CONTAINERS: their own forward is never
called on the GPU path.
"""
import torch
import torch.nn as nn


def conv3x3(in_planes, out_planes, stride=1):
    """3x3 convolution with padding (reference backbone.py:20-23)."""
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


def conv1x1(in_planes, out_planes, stride=1):
    """1x1 convolution (reference backbone.py:26-28)."""
    return nn.Conv2d(in_planes, out_planes, kernel_size=1, stride=stride, bias=False)


class BasicBlock(nn.Module):
    """Parameter layout of reference backbone.py:31-50."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):  # pragma: no cover - blocks are executed by the EncoderEngine
        raise RuntimeError("gdl_b200 BasicBlock is a parameter container; run the enclosing ResNet")


class ResNet(nn.Module):
    """reference backbone.py:73-201 (no avgpool / fc; audio Cin=1, visual Cin=3)."""

    def __init__(self, args, block, layers, modality):
        super().__init__()
        self.modality = modality
        self.inplanes = 64
        if modality == 'audio':
            self.conv1 = nn.Conv2d(1, self.inplanes, kernel_size=7, stride=2, padding=3, bias=False)
        elif modality == 'visual':
            self.conv1 = nn.Conv2d(3, self.inplanes, kernel_size=7, stride=2, padding=3, bias=False)
        else:
            raise NotImplementedError(
                'Incorrect modality, should be audio or visual but got {}'.format(modality))
        self.bn1 = nn.BatchNorm2d(self.inplanes)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2)
        self.args = args
        # reference backbone.py:117-122
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
                nn.init.normal_(m.weight, mean=1, std=0.02)
                nn.init.constant_(m.bias, 0)
        self._engines = {}

    def _make_layer(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(conv1x1(self.inplanes, planes * block.expansion, stride),
                                       nn.BatchNorm2d(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes))
        return nn.Sequential(*layers)

    # ------------------------------------------------------------------ engine plumbing
    def engine(self, N, H, W):
        """EncoderEngine for N images of H x W on this module's device (cached per geometry)."""
        from .engine import EncoderEngine
        dev = self.conv1.weight.device
        key = (N, H, W, dev)
        eng = self._engines.get(key)
        if eng is None:
            if dev.type != 'cuda':
                raise RuntimeError("gdl_b200.ResNet runs on a B200 only (no CPU fallback); "
                                   "move the model to cuda first")
            eng = EncoderEngine(self, N, H, W, dev)
            # two live geometries (the full batch and an epoch's short tail batch): activation arenas are large
            if len(self._engines) >= 2:
                self._engines.pop(next(iter(self._engines)))
        else:
            self._engines.pop(key)
        self._engines[key] = eng  # most recently used last
        return eng

    def forward(self, x):
        """Returns the layer4 feature map as fp32 NCHW like the reference (backbone.py:160-201):
        audio [B,1,F,T] -> [B,512,h,w]; visual [B,3,T,H,W] -> [B*T,512,7,7]."""
        from .autograd import encoder_map
        return encoder_map(self, x)


def resnet18(modality, args, progress=True, **kwargs):
    return ResNet(args, BasicBlock, [2, 2, 2, 2], modality)
