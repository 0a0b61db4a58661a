"""Layer geometry of the networks on the hot path.

Shapes follow the reference's extern weight declarations for the BLAS build
(Network.cpp:82-107: NN128 policy stack) and the value net (Network.cpp:110-137), in the order
Network::initialize pushes them (Network.cpp:206-233). Every conv is followed by bias + ELU
(Network.cpp:382-392); val_ip13 has ELU, val_ip14 is linear (Network.cpp:411-421).
"""
from __future__ import annotations

import dataclasses

P = 361            # 19 x 19 board points
BOARD = 19
POLICY, VALUE = 0, 1


@dataclasses.dataclass(frozen=True)
class Conv:
    k: int
    c_in: int
    c_out: int

    @property
    def n_weights(self) -> int:
        return self.k * self.k * self.c_in * self.c_out

    @property
    def fan_in(self) -> int:
        return self.k * self.k * self.c_in


@dataclasses.dataclass(frozen=True)
class InnerProduct:
    n_in: int
    n_out: int


POLICY_CONVS = (Conv(5, 32, 96), Conv(3, 96, 128)) + (Conv(3, 128, 128),) * 10 + (Conv(3, 128, 1),)
# the OpenCL build's policy net (Network.cpp:55-80): 5x5 32->128, 3x3 128->192, 10x 192->192, 192->1
POLICY192_CONVS = (Conv(5, 32, 128), Conv(3, 128, 192)) + (Conv(3, 192, 192),) * 10 + (Conv(3, 192, 1),)
VALUE_CONVS = (Conv(5, 32, 64),) + (Conv(3, 64, 64),) * 10 + (Conv(3, 64, 1),)
VALUE_IPS = (InnerProduct(361, 256), InnerProduct(256, 1))

# Dense im2col-GEMM FLOPs per position (2 x MACs), exactly what the reference's cblas_sgemm
# executes (SURVEY.md section 8d / BASELINE.md section 3).
POLICY_FLOPS = 2 * P * sum(c.k * c.k * c.c_in * c.c_out for c in POLICY_CONVS)
VALUE_FLOPS = 2 * P * sum(c.k * c.k * c.c_in * c.c_out for c in VALUE_CONVS) \
    + 2 * sum(i.n_in * i.n_out for i in VALUE_IPS)
POLICY192_FLOPS = 2 * P * sum(c.k * c.k * c.c_in * c.c_out for c in POLICY192_CONVS)
assert POLICY_FLOPS == 1_200_761_088 and VALUE_FLOPS == 303_725_696 and POLICY192_FLOPS == 2_630_297_984
