import torch.nn as nn


class SeqToANNContainer(nn.Sequential):
    """Flatten [T, B] -> T*B, apply, un-flatten (spikingjelly layer.SeqToANNContainer)."""

    def forward(self, x_seq):
        head = [x_seq.shape[0], x_seq.shape[1]]
        y = super().forward(x_seq.flatten(0, 1))
        return y.view(head + list(y.shape[1:]))


class BatchNorm2d(nn.BatchNorm2d):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True,
                 step_mode="s"):
        super().__init__(num_features, eps, momentum, affine, track_running_stats)
        self.step_mode = step_mode

    def forward(self, x):
        if self.step_mode == "s":
            return super().forward(x)
        if x.dim() != 5:
            raise ValueError("expected [T, N, C, H, W]")
        y = super().forward(x.flatten(0, 1))
        return y.view([x.shape[0], x.shape[1]] + list(y.shape[1:]))
