"""ResNet image encoders of 3D-FM GAN (E_Tsr: ``resnet18(tensor_encoding=True)`` -> [B,512,4,4];
E_W: ``resnet18(tensor_encoding=False)`` -> [B,512]).

Module / state-dict layout mirrors the reference's ``resnet_encoder.py`` (a torchvision ResNet
with the classifier removed, :152-283) so checkpoints are interchangeable.  In eval mode on CUDA
the forward runs on the tcgen05 implicit-GEMM engine with BatchNorm, ReLU and the residual add
folded into the convolution epilogues (``fm3d.encoder_engine``); otherwise (training) it is the
plain differentiable PyTorch composition.
"""
import os

import torch
from torch import nn


def conv3x3(in_planes, out_planes, stride=1, groups=1, dilation=1):
    return nn.Conv2d(in_planes, out_planes, 3, stride=stride, padding=dilation, groups=groups, bias=False,
                     dilation=dilation)


def conv1x1(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, 1, stride=stride, bias=False)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=1, base_width=64, dilation=1,
                 norm_layer=None):
        super().__init__()
        norm_layer = norm_layer or nn.BatchNorm2d
        if groups != 1 or base_width != 64:
            raise ValueError('BasicBlock only supports groups=1 and base_width=64')
        if dilation > 1:
            raise NotImplementedError('Dilation > 1 not supported in BasicBlock')
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.bn1 = norm_layer(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = norm_layer(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        shortcut = x if self.downsample is None else self.downsample(x)
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return self.relu(y + shortcut)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=1, base_width=64, dilation=1,
                 norm_layer=None):
        super().__init__()
        norm_layer = norm_layer or nn.BatchNorm2d
        width = int(planes * (base_width / 64.)) * groups
        self.conv1 = conv1x1(inplanes, width)
        self.bn1 = norm_layer(width)
        self.conv2 = conv3x3(width, width, stride, groups, dilation)
        self.bn2 = norm_layer(width)
        self.conv3 = conv1x1(width, planes * self.expansion)
        self.bn3 = norm_layer(planes * self.expansion)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        shortcut = x if self.downsample is None else self.downsample(x)
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return self.relu(y + shortcut)


class ResNet(nn.Module):
    def __init__(self, block, layers, num_classes=1000, zero_init_residual=False, groups=1, width_per_group=64,
                 replace_stride_with_dilation=None, norm_layer=None, tensor_encoding=True, tensor_transform=False):
        """tensor_encoding: True -> [B,512,4,4] feature tensor (AvgPool 2x2), False -> [B,512] vector;
        tensor_transform: additionally map the flattened tensor to a 512-vector (2-encoder scheme)."""
        super().__init__()
        self._norm_layer = norm_layer or nn.BatchNorm2d
        self.tensor_encoding = tensor_encoding
        self.tensor_transform = tensor_transform
        self.inplanes = 64
        self.dilation = 1
        if replace_stride_with_dilation is None:
            replace_stride_with_dilation = [False, False, False]
        if len(replace_stride_with_dilation) != 3:
            raise ValueError('replace_stride_with_dilation should be None or a 3-element tuple, got '
                             f'{replace_stride_with_dilation}')
        self.groups = groups
        self.base_width = width_per_group
        self.conv1 = nn.Conv2d(3, self.inplanes, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = self._norm_layer(self.inplanes)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2, dilate=replace_stride_with_dilation[0])
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2, dilate=replace_stride_with_dilation[1])
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2, dilate=replace_stride_with_dilation[2])
        self.avgpool = nn.AvgPool2d(kernel_size=2, stride=2) if tensor_encoding else nn.AdaptiveAvgPool2d((1, 1))
        if tensor_transform:
            self.ten_fc = nn.Linear(512 * 4 * 4, 512)

        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        if zero_init_residual:
            for m in self.modules():
                if isinstance(m, Bottleneck):
                    nn.init.constant_(m.bn3.weight, 0)
                elif isinstance(m, BasicBlock):
                    nn.init.constant_(m.bn2.weight, 0)
        if os.environ.get("FM3D_NATIVE_ENC", "1") != "0":
            # under autograd (training) the convolutions differentiate through the tcgen05 kernels, like G's and D's
            from fm3d.convgrad import use_native_convs
            use_native_convs(self)

    def _make_layer(self, block, planes, blocks, stride=1, dilate=False):
        prev_dilation = self.dilation
        if dilate:
            self.dilation *= stride
            stride = 1
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(conv1x1(self.inplanes, planes * block.expansion, stride),
                                       self._norm_layer(planes * block.expansion))
        stack = [block(self.inplanes, planes, stride, downsample, self.groups, self.base_width, prev_dilation,
                       self._norm_layer)]
        self.inplanes = planes * block.expansion
        stack += [block(self.inplanes, planes, groups=self.groups, base_width=self.base_width,
                        dilation=self.dilation, norm_layer=self._norm_layer) for _ in range(1, blocks)]
        return nn.Sequential(*stack)

    def _engine_ok(self, x):
        return (x.is_cuda and not self.training and not self.tensor_transform and
                os.environ.get("FM3D_ENGINE", "1") != "0" and
                not (torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))))

    def _forward_impl(self, x):
        if self._engine_ok(x):
            from fm3d.encoder_engine import run_resnet
            out = run_resnet(self, x)
            if out is not None:
                return out
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        x = self.avgpool(x)
        if not self.tensor_encoding:
            x = torch.flatten(x, 1)
        if self.tensor_transform:
            return x, self.ten_fc(torch.flatten(x, 1))
        return x

    def forward(self, x):
        return self._forward_impl(x)


def _resnet(block, layers, pretrained, **kwargs):
    if pretrained:
        raise RuntimeError('pretrained ImageNet weights are not available offline; load a state dict instead')
    return ResNet(block, layers, **kwargs)


def resnet18(pretrained=False, progress=True, **kwargs):
    return _resnet(BasicBlock, [2, 2, 2, 2], pretrained, **kwargs)


def resnet34(pretrained=False, progress=True, **kwargs):
    return _resnet(BasicBlock, [3, 4, 6, 3], pretrained, **kwargs)


def resnet50(pretrained=False, progress=True, **kwargs):
    return _resnet(Bottleneck, [3, 4, 6, 3], pretrained, **kwargs)
