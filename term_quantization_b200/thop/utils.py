from collections.abc import Iterable


def clever_format(nums, format="%.2f"):
    """1234567 -> '1.23M' (thop/utils.py:4-23); a single number in, a single string out."""
    if not isinstance(nums, Iterable):
        nums = [nums]
    out = []
    for num in nums:
        for limit, suffix in ((1e12, "T"), (1e9, "G"), (1e6, "M"), (1e3, "K")):
            if num > limit:
                out.append(format % (num / limit) + suffix)
                break
        else:
            out.append(format % num + "B")
    return out[0] if len(out) == 1 else tuple(out)
