"""What the trunk launch costs in the shifted-row formulation, from the layer shapes alone (no GPU): MMA instructions per work item,
their ideal cost on the tensor pipe and the cost the shared-memory port allows, and the launch time that follows at a given SM clock.
DESIGN.md section 5 quotes these figures beside the measured launch.

  python tools/trunk_model.py [--batch 256] [--mhz 1650] [--policy-precision 0] [--value-precision 1] [--measured-us 369]

Model: a work item is a CTA pair's 512 rows (two 256-row tiles); per 16-channel slab, tap and K-loop term it issues TWO
tcgen05.mma.cta_group::2 (row halves), M = 256 over the pair, N = c_out, K = 16 (kind::f16) or 32 (kind::f8f6f4).
  tensor pipe: 8192 dense 16-bit flop / clk / SM  ->  an M = 128-per-SM x N x 16 instruction takes N / 2 cycles
  shared memory: 128 B / clk / SM; per instruction an SM reads its A tile (128 rows x 32 B = 4 KB) and serves its half of B
  (N / 2 rows x 32 B) to both tensor cores of the pair -> (4096 + N * 32) / 128 cycles
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from leela_b200 import netdefs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--mhz", type=float, default=1650.0, help="SM clock inside the kernel (tools/trace_timeline.py prints it)")
    ap.add_argument("--clusters", type=int, default=74)
    ap.add_argument("--policy-precision", type=int, default=0)
    ap.add_argument("--value-precision", type=int, default=1)
    ap.add_argument("--measured-us", type=float, default=None)
    a = ap.parse_args()
    tot_ideal = tot_smem = 0.0
    algo = 0
    print(f"{'net':7s}{'layer':>6s}{'shape':>14s}{'items':>7s}{'MMAs/item':>10s}{'ideal cyc':>10s}{'smem cyc':>9s}{'bound cyc/item':>15s}")
    for name, convs, mode in (("policy", netdefs.POLICY_CONVS, a.policy_precision), ("value", netdefs.VALUE_CONVS, a.value_precision)):
        for i, c in enumerate(convs[:-1]):
            rows = a.batch * (441 if c.k == 5 else 400)
            items = -(-rows // 512)
            terms = {0: 1, 1: 2, 2: 2 if i == 0 else 3}[mode]
            mmas = (c.c_in // 16) * c.k * c.k * 2 * terms
            ideal = c.c_out / 2
            smem = (4096 + c.c_out * 32) / 128
            per_item = mmas * max(ideal, smem)
            tot_ideal += items * mmas * ideal
            tot_smem += items * per_item
            algo += 2 * 361 * a.batch * c.k * c.k * c.c_in * c.c_out
            if i < 3 or i == len(convs) - 2:
                print(f"{name:7s}{i + 1:6d}{f'{c.k}x{c.k} {c.c_in}->{c.c_out}':>14s}{items:7d}{mmas:10d}{ideal:10.0f}{smem:9.0f}{per_item:15.0f}")
            elif i == 3:
                print(f"{name:7s}   ...")
    us = lambda cyc: cyc / a.clusters / a.mhz
    print(f"\ncluster-cycles per launch: {tot_ideal / 1e6:.1f} M on the tensor pipe alone, {tot_smem / 1e6:.1f} M with the shared-memory bound")
    print(f"launch at {a.mhz:.0f} MHz on {a.clusters} clusters: {us(tot_ideal):.0f} us (tensor pipe), {us(tot_smem):.0f} us (formulation bound)")
    print(f"algorithmic flops per launch {algo / 1e9:.1f} G; executed (issued) {tot_ideal * 2 * 8192 / 1e9:.1f} G bf16-equivalent = x{tot_ideal * 2 * 8192 / algo:.3f}")
    if a.measured_us:
        print(f"measured {a.measured_us:.1f} us: {us(tot_smem) / a.measured_us:.2f} of the formulation bound, tensor pipe busy {us(tot_ideal) / a.measured_us:.2f}")


if __name__ == "__main__":
    main()
