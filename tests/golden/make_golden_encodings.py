"""Generates tests/golden/enc_golden.npz: reference-held pins for the two ADDED encodings
(TQ_ENC_BINARY, TQ_ENC_BOOTH; the reference kernel itself implements HESE only).

Run only where /root/reference exists:  python tests/golden/make_golden_encodings.py
The .npz is committed; tests read it and never need /root/reference.

BINARY -- bit_utils.expand_binary_bits (bit_utils.py:63-73), the reference's own binary term expansion
  (q = floor(W / sf + 0.5), MSB-first bit planes of |q|).  bit_utils imports cleanly but the function ends in
  `.cuda()`; only that call is stripped from the function's text before it is executed here.
BOOTH  -- the truth table of verilog/booth_encoder.v:57-78 (radix-2: {delayed, current} = 01 -> +1, 10 -> -1,
  00 / 11 -> 0), clocked bit-serially exactly as the module's two input registers see the stream.  The stream is
  fed MSB first behind a leading 0 and followed by a trailing 0 (the register reset value / pipeline flush): the
  only order for which the emitted digits reproduce the input value (checked below for every q).
"""
import os
import re
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"


def reference_expand_binary_bits():
    src = open(os.path.join(REF, "bit_utils.py")).read()
    fn = re.search(r"^def expand_binary_bits\(W, sf, min_bits=9\):.*?^    return W\s*\Z", src, re.S | re.M).group(0)
    assert fn.count(".cuda()") == 1
    ns = {"torch": torch}
    exec(fn.replace(".cuda()", ""), ns)
    return ns["expand_binary_bits"]


def booth_table():
    """case ({dummy[1], dummy[0]}) of verilog/booth_encoder.v:57-78 -> (output_stream, sign_stream)."""
    src = open(os.path.join(REF, "verilog", "booth_encoder.v")).read()
    table = {}
    for key, body in re.findall(r"2'b([01]{2})\s*:\s*begin(.*?)end", src, re.S):
        out = int(re.search(r"output_stream_reg\s*<=\s*'b([01])", body).group(1))
        sgn = int(re.search(r"sign_stream_reg\s*<=\s*'b([01])", body).group(1))
        table[key] = (out, sgn)
    assert table == {"00": (0, 0), "01": (1, 0), "10": (1, 1), "11": (0, 0)}, table
    return table


def booth_digits(q, nbits, table):
    """Clock the encoder over q's bits, MSB first: input_stream_reg <= bit, input_stream_reg_delay <= previous
    (booth_encoder.v:38-45); the digit emitted while {delayed, current} = (b[j+1], b[j]) has weight 2^(j+1)."""
    stream = [(q >> j) & 1 for j in range(nbits - 1, -1, -1)] + [0]      # trailing 0 flushes the LSB's digit
    delayed, P, N = 0, 0, 0
    for step, cur in enumerate(stream):
        out, sgn = table[f"{delayed}{cur}"]
        weight = nbits - step                                              # bit position of this digit
        if out:
            if sgn:
                N |= 1 << weight
            else:
                P |= 1 << weight
        delayed = cur
    return P, N


def main():
    rng = np.random.default_rng(7)
    expand = reference_expand_binary_bits()
    # values kept away from rounding ties of floor(W/sf + 0.5) for negative W (the kernel rounds |W| half up and
    # applies the sign; bit_utils takes |.| AFTER the floor, which differs exactly on negative ties)
    W = np.concatenate([rng.uniform(-1, 1, 4000), [0.0, 1.0, -1.0, 0.999, 0.5, 0.25, 0.0019, 0.002]]).astype(np.float32)
    bits = 9
    sf = np.float32(1.0 / 2 ** (bits - 1))
    q = np.floor(W.astype(np.float64) / float(sf) + 0.5)
    keep = (np.abs(W.astype(np.float64) / float(sf) - np.round(W.astype(np.float64) / float(sf))) > 1e-3) | (W >= 0)
    W = W[keep & (np.abs(q) < 2 ** bits)]
    planes = expand(torch.from_numpy(W), float(sf), bits).numpy().astype(np.uint8)      # [n, bits], MSB first

    table = booth_table()
    nb = 12
    qs = np.arange(0, 1 << nb, dtype=np.int64)
    P = np.zeros_like(qs)
    N = np.zeros_like(qs)
    for i, v in enumerate(qs):
        P[i], N[i] = booth_digits(int(v), nb, table)
        assert P[i] - N[i] == v and P[i] & N[i] == 0
    out = os.path.join(ROOT, "tests", "golden", "enc_golden.npz")
    np.savez_compressed(out, bin_W=W, bin_sf=sf, bin_bits=bits, bin_planes=planes, booth_q=qs, booth_P=P, booth_N=N)
    print("wrote", out, "binary cases", len(W), "booth cases", len(qs))


if __name__ == "__main__":
    main()
