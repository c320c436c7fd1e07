"""CPU oracle (TEST INFRASTRUCTURE ONLY): PyTorch fp32 restatement of the model graphs the
reference executes as an opaque TensorRT engine (src/tensorrt/img2img_infer.cpp:80).

The graphs are third-party (nagadomi/nunif waifu2x `cunet` / `swin_unet`, README.md:99),
fetched by the reference as ONNX files at run time (src/main.cpp:201-204) and absent from
/root/reference and from this image: PARITY UNPINNED.  The arithmetic below
follows the published nunif model definitions; weights are seeded synthetic.

Never imported by the product path.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class SEBlock(nn.Module):
    """Squeeze/excite: mean_{H,W} -> 1x1 conv C->C/r -> ReLU -> 1x1 conv -> sigmoid -> scale."""

    def __init__(self, channels: int, reduction: int = 8):
        super().__init__()
        self.conv1 = nn.Conv2d(channels, channels // reduction, 1, 1, 0, bias=True)
        self.conv2 = nn.Conv2d(channels // reduction, channels, 1, 1, 0, bias=True)

    def forward(self, x):
        z = F.adaptive_avg_pool2d(x, 1)
        z = F.relu(self.conv1(z))
        z = torch.sigmoid(self.conv2(z))
        return x * z


class UNetConv(nn.Module):
    def __init__(self, cin: int, mid: int, cout: int, se: bool):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(cin, mid, 3, 1, 0), nn.LeakyReLU(0.1),
            nn.Conv2d(mid, cout, 3, 1, 0), nn.LeakyReLU(0.1))
        self.se = SEBlock(cout, 8) if se else None

    def forward(self, x):
        z = self.conv(x)
        return self.se(z) if self.se is not None else z


class UNet1(nn.Module):
    def __init__(self, cin: int, cout: int, deconv: bool):
        super().__init__()
        self.conv1 = UNetConv(cin, 32, 64, se=False)
        self.conv1_down = nn.Conv2d(64, 64, 2, 2, 0)
        self.conv2 = UNetConv(64, 128, 64, se=True)
        self.conv2_up = nn.ConvTranspose2d(64, 64, 2, 2, 0)
        self.conv3 = nn.Conv2d(64, 64, 3, 1, 0)
        self.conv_bottom = nn.ConvTranspose2d(64, cout, 4, 2, 3) if deconv else nn.Conv2d(64, cout, 3, 1, 0)

    def forward(self, x):
        x1 = self.conv1(x)
        x2 = F.leaky_relu(self.conv1_down(x1), 0.1)
        x2 = self.conv2(x2)
        x2 = F.leaky_relu(self.conv2_up(x2), 0.1)
        x1 = F.pad(x1, (-4, -4, -4, -4))
        x3 = F.leaky_relu(self.conv3(x1 + x2), 0.1)
        return self.conv_bottom(x3)


class UNet2(nn.Module):
    def __init__(self, cin: int, cout: int, deconv: bool):
        super().__init__()
        self.conv1 = UNetConv(cin, 32, 64, se=False)
        self.conv1_down = nn.Conv2d(64, 64, 2, 2, 0)
        self.conv2 = UNetConv(64, 64, 128, se=True)
        self.conv2_down = nn.Conv2d(128, 128, 2, 2, 0)
        self.conv3 = UNetConv(128, 256, 128, se=True)
        self.conv3_up = nn.ConvTranspose2d(128, 128, 2, 2, 0)
        self.conv4 = UNetConv(128, 64, 64, se=True)
        self.conv4_up = nn.ConvTranspose2d(64, 64, 2, 2, 0)
        self.conv5 = nn.Conv2d(64, 64, 3, 1, 0)
        self.conv_bottom = nn.ConvTranspose2d(64, cout, 4, 2, 3) if deconv else nn.Conv2d(64, cout, 3, 1, 0)

    def forward(self, x):
        x1 = self.conv1(x)
        x2 = F.leaky_relu(self.conv1_down(x1), 0.1)
        x2 = self.conv2(x2)
        x3 = F.leaky_relu(self.conv2_down(x2), 0.1)
        x3 = self.conv3(x3)
        x3 = F.leaky_relu(self.conv3_up(x3), 0.1)
        x2 = F.pad(x2, (-4, -4, -4, -4))
        x4 = self.conv4(x2 + x3)
        x4 = F.leaky_relu(self.conv4_up(x4), 0.1)
        x1 = F.pad(x1, (-16, -16, -16, -16))
        x5 = F.leaky_relu(self.conv5(x1 + x4), 0.1)
        return self.conv_bottom(x5)


class CUNet(nn.Module):
    """cunet/art scale 1 (denoise): offset 28, out = T - 56."""
    scale = 1
    offset = 28

    def __init__(self):
        super().__init__()
        self.unet1 = UNet1(3, 3, deconv=False)
        self.unet2 = UNet2(3, 3, deconv=False)

    def forward(self, x):
        z1 = self.unet1(x)
        z2 = self.unet2(z1)
        z1 = F.pad(z1, (-20, -20, -20, -20))
        return torch.clamp(z1 + z2, 0.0, 1.0)


class UpCUNet(nn.Module):
    """cunet/art scale 2: offset 36 (in output pixels), out = 2*T - 72."""
    scale = 2
    offset = 36

    def __init__(self):
        super().__init__()
        self.unet1 = UNet1(3, 3, deconv=True)
        self.unet2 = UNet2(3, 3, deconv=False)

    def forward(self, x):
        z1 = self.unet1(x)
        z2 = self.unet2(z1)
        z1 = F.pad(z1, (-20, -20, -20, -20))
        return torch.clamp(z1 + z2, 0.0, 1.0)


def cunet_out_size(scale: int, tile: int) -> int:
    return tile - 56 if scale == 1 else 2 * tile - 72


def synth_init_(model: nn.Module, seed: int = 1234) -> nn.Module:
    """Seeded synthetic weights: variance-preserving for LeakyReLU(0.1), the image
    heads centred on 0.5 so u8 outputs populate [0,255] and the +-1 LSB test is meaningful."""
    g = torch.Generator().manual_seed(seed)
    gain = math.sqrt(2.0 / (1.0 + 0.1 ** 2))
    for name, m in model.named_modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            w = m.weight
            if isinstance(m, nn.ConvTranspose2d):
                # each output pixel sums over cin * (k/stride)^2 taps
                fan = w.shape[0] * (w.shape[2] // m.stride[0]) * (w.shape[3] // m.stride[1])
            else:
                fan = w.shape[1] * w.shape[2] * w.shape[3]
            std = gain / math.sqrt(fan)
            if ".se." in name:
                std = 1.0 / math.sqrt(fan)
            if name.endswith("conv_bottom"):
                std = 0.25 / math.sqrt(fan)
            with torch.no_grad():
                w.copy_(torch.randn(w.shape, generator=g) * std)
                if m.bias is not None:
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.05)
                    if name == "unet1.conv_bottom":
                        m.bias.add_(0.5)
    # first layers see inputs in [0,1] (mean 0.5): centre them so activations are not all-positive
    for first in ("unet1.conv1.conv.0", "unet2.conv1.conv.0"):
        mods = dict(model.named_modules())
        if first in mods:
            m = mods[first]
            with torch.no_grad():
                m.weight.mul_(3.0)
                m.bias.sub_(0.5 * m.weight.sum(dim=(1, 2, 3)))
    return model.eval()


def make_model(family: str, scale: int, seed: int = 1234) -> nn.Module:
    if family == "swin_unet":
        return synth_init_swin_(SwinUNet(scale), seed)
    if family == "cunet":
        if scale == 1:
            return synth_init_(CUNet(), seed)
        if scale == 2:
            return synth_init_(UpCUNet(), seed)
        raise ValueError("cunet/art does not support scale factor 4.")  # main.cpp:142-143
    raise ValueError(family)


# ------------------------------------------------------------------------------------------------------------------
# SwinUNet family (swin_unet/{art,art_scan,photo}): nunif `waifu2x/models/swin_unet.py` restated.
# Blocks are torchvision's SwinTransformerBlock v1 (pre-LN, W-MSA / SW-MSA with relative position bias, GELU MLP).
# ------------------------------------------------------------------------------------------------------------------
class SwinBlocks(nn.Module):
    def __init__(self, dim: int, heads: int, layers: int, window: int = 6, mlp_ratio: float = 2.0):
        super().__init__()
        from torchvision.models.swin_transformer import SwinTransformerBlock
        self.block = nn.Sequential(*[
            SwinTransformerBlock(dim, heads, window_size=[window, window],
                                 shift_size=[0 if i % 2 == 0 else window // 2] * 2, mlp_ratio=mlp_ratio,
                                 dropout=0.0, attention_dropout=0.0, stochastic_depth_prob=0.0, norm_layer=nn.LayerNorm)
            for i in range(layers)])

    def forward(self, x):  # BHWC
        return self.block(x)


class PatchDown(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 2, 2, 0)

    def forward(self, x):  # BHWC
        return self.conv(x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)


class PatchUp(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.proj = nn.Linear(cin, cout * 4)

    def forward(self, x):  # BHWC
        x = self.proj(x).permute(0, 3, 1, 2)
        return F.pixel_shuffle(x, 2).permute(0, 2, 3, 1)


class ToImage(nn.Module):
    def __init__(self, cin: int, cout: int, scale: int):
        super().__init__()
        self.scale = scale
        self.proj = nn.Linear(cin, cout * scale * scale)

    def forward(self, x):  # BHWC -> BCHW
        x = self.proj(x).permute(0, 3, 1, 2)
        return F.pixel_shuffle(x, self.scale) if self.scale > 1 else x


class SwinUNet(nn.Module):
    """offset = 8 * scale output pixels per side; out = (T - 16) * scale; needs (T - 16) % 48 == 0."""

    def __init__(self, scale: int = 4, base_dim: int = 96, base_layers: int = 2, window: int = 6):
        super().__init__()
        assert scale in (1, 2, 4)
        C, H, L = base_dim, base_dim // 16, base_layers
        self.scale = scale
        self.offset = 8 * scale
        self.patch = nn.Sequential(nn.Conv2d(3, C // 2, 3, 1, 0), nn.LeakyReLU(0.1), nn.Conv2d(C // 2, C, 3, 1, 0), nn.LeakyReLU(0.1))
        self.swin1 = SwinBlocks(C, H, L, window)
        self.down1 = PatchDown(C, C * 2)
        self.swin2 = SwinBlocks(C * 2, H, L, window)
        self.down2 = PatchDown(C * 2, C * 2)
        self.swin3 = SwinBlocks(C * 2, H, L * 3, window)
        self.up2 = PatchUp(C * 2, C * 2)
        self.swin4 = SwinBlocks(C * 2, H, L, window)
        self.up1 = PatchUp(C * 2, C)
        self.swin5 = SwinBlocks(C, H, L, window)
        if scale == 4:
            self.up0 = PatchUp(C, C)
            self.to_image = ToImage(C, 3, 2)
        else:
            self.up0 = None
            self.to_image = ToImage(C, 3, scale)

    def forward(self, x):
        x2 = F.pad(self.patch(x), (-6, -6, -6, -6)).permute(0, 2, 3, 1)  # BHWC
        x3 = self.swin1(x2)
        x4 = self.swin2(self.down1(x3))
        x5 = self.swin3(self.down2(x4))
        x = self.swin4(self.up2(x5) + x4)
        x = self.swin5(self.up1(x) + x3)
        if self.up0 is not None:
            x = self.up0(x)
        return torch.clamp(self.to_image(x), 0.0, 1.0)


def swin_out_size(scale: int, tile: int) -> int:
    return (tile - 16) * scale


def synth_init_swin_(model: nn.Module, seed: int = 1234) -> nn.Module:
    """Seeded synthetic SwinUNet weights: unit-variance token stream, small residual branches, image head centred at 0.5."""
    g = torch.Generator().manual_seed(seed)
    gain = math.sqrt(2.0 / (1.0 + 0.1 ** 2))
    with torch.no_grad():
        for name, m in model.named_modules():
            if isinstance(m, nn.Conv2d):
                fan = m.weight.shape[1] * m.weight.shape[2] * m.weight.shape[3]
                std = (gain if name.startswith("patch") else 1.0) / math.sqrt(fan)
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * std)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.05)
            elif isinstance(m, nn.Linear):
                fan = m.weight.shape[1]
                std = 1.0 / math.sqrt(fan)
                if name.endswith("attn.proj") or name.endswith("mlp.3"):
                    std *= 0.5  # residual branches
                if name == "to_image.proj":
                    std *= 0.2
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * std)
                if m.bias is not None:
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.02)
                    if name == "to_image.proj":
                        m.bias.add_(0.5)
            elif isinstance(m, nn.LayerNorm):
                m.weight.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.05 * torch.randn(m.bias.shape, generator=g))
        for name, p in model.named_parameters():
            if name.endswith("relative_position_bias_table"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.5)
        m = model.patch[0]
        m.weight.mul_(3.0)
        m.bias.sub_(0.5 * m.weight.sum(dim=(1, 2, 3)))
    return model.eval()
