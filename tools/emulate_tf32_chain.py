"""CPU emulation of the 3xTF32 GEMM accumulation of csrc/linear_tf32.cu with a TRUNCATING accumulator (what the tensor cores
do), to size the effect of the chain length:  python tools/emulate_tf32_chain.py [K]

Result for the failing parity case (R=129, K=768): one mma chain over all of K -> 6.7e-6 of max|y| (the B200 run left the
2e-6 band); each 32-wide K chunk summed from zero and added with a rounded fp32 add -> 4.1e-7."""
import sys

import numpy as np
import torch


def rz32(x64):
    y = x64.astype(np.float32)
    bad = np.abs(y.astype(np.float64)) > np.abs(x64)
    y[bad] = np.nextafter(y[bad], np.float32(0))
    return y


def tf32(v):
    i = v.view(np.int32)
    return ((i + 0x1000) & ~0x1fff).astype(np.int32).view(np.float32)


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 768
    R, N = 129, 256
    g = torch.Generator().manual_seed(R + K + 2304)
    x = torch.randn(R, K, generator=g).numpy()
    w = (torch.randn(2304, K, generator=g) * K ** -0.5).numpy()[:N]
    ref = x.astype(np.float64) @ w.astype(np.float64).T
    xh, wh = tf32(x.copy()), tf32(w.copy())
    xl, wl = tf32(x - xh), tf32(w - wh)

    def run(flush_every):
        acc = np.zeros((R, N), np.float32)
        part = np.zeros((R, N), np.float32)
        steps = 0
        for k in range(0, K, 8):
            s = slice(k, k + 8)
            for a, b in ((xl, wh), (xh, wl), (xh, wh)):                      # the three mma of one k8 step, small terms first
                prod = a[:, s].astype(np.float64) @ b[:, s].astype(np.float64).T
                part = rz32(part.astype(np.float64) + prod)                  # the accumulator update truncates
            steps += 1
            if flush_every and steps % flush_every == 0:
                acc = (acc + part).astype(np.float32)                        # rounded fp32 add outside the tensor cores
                part[:] = 0
        acc = part if not flush_every else (acc + part).astype(np.float32)
        return float(np.abs(acc.astype(np.float64) - ref).max() / np.abs(ref).max())

    print(f"K={K}: one chain {run(0):.2e}   flush every 32 columns {run(4):.2e}")


if __name__ == "__main__":
    main()
