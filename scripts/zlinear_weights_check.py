"""Groundwork for the next addend kernel: along a z-line of the dense grid the left trilinear weight
w0(s) of a voxel level is linear in the step s inside a voxel cell (slope -(R-1)/(res-1), 0 where the coordinate is
clamped by border padding), and the slope is shared by the three W-shift classes of a level.  A kernel can therefore
keep, per channel, P = sum_c (G1_c - a_c D_c) and Q = -sum_c b_c D_c and evaluate P + s*Q with ONE FMA per step instead
of one per class.  This script bounds what that costs numerically: the largest difference between the exact fp32
weights (ATen's formula, as csrc/common.cuh axis_border computes them) and the per-segment linear model.
Result (res 64..512, R 8..32, all three shifts): 3.7e-6 -- three orders below the bf16 rounding of the addend."""
import numpy as np

f32 = np.float32


def linspace_q(res):
    i = np.arange(res, dtype=np.float64)
    v = i * (1.0 / (res - 1)) - 0.5
    v[-1] = 0.5
    return v.astype(f32) * f32(2.0)


def axis_border(c, R):
    i = ((c + f32(1.0)) * f32(0.5)) * f32(R - 1)
    i = np.minimum(np.maximum(i, f32(0)), f32(R - 1)).astype(f32)
    f = np.floor(i)
    return f.astype(int), ((f + f32(1.0)) - i).astype(f32)


def main():
    worst = 0.0
    for res in (64, 128, 256, 512):
        q = linspace_q(res)
        for R in (8, 16, 32):
            for sh in (0.0, -0.0722, 0.0722):
                c = (q + f32(sh)).astype(f32) if sh else q
                i0, w0 = axis_border(c, R)
                raw = ((c + f32(1.0)) * f32(0.5)) * f32(R - 1)
                state = np.where(raw < 0, -1, np.where(raw > R - 1, 1, 0))
                key = i0 * 4 + state + 1
                s = 0
                while s < res:
                    e = s
                    while e + 1 < res and key[e + 1] == key[s]:
                        e += 1
                    slope = 0.0 if state[s] != 0 else -((R - 1) / (res - 1))
                    model = np.float64(w0[s]) + slope * np.arange(e - s + 1, dtype=np.float64)
                    worst = max(worst, float(np.abs(model - w0[s:e + 1].astype(np.float64)).max()))
                    s = e + 1
    print(f"max |linear model - exact w0| over all segments: {worst:.3e}")


if __name__ == "__main__":
    main()
