"""numpy walk over the PACKED arrays (tests only): proves on the CPU that the layout the packer emits carries
exactly the information the kernels need -- it mirrors k_weights_m4 / k_column_reduce / k_locus_acc /
k_locus_update of gbrs_b200/csrc/em_kernels.cu one to one, without their parallel decomposition."""
import numpy as np


def unpack_entries(ent, entry_bytes):
    sh = 8 * entry_bytes - 8
    ent = ent.astype(np.uint64)
    return (ent & np.uint64((1 << sh) - 1)).astype(np.int64), (ent >> np.uint64(sh)).astype(np.int64)


def deinterleave(a, item_off, item_len=64):
    """Undo the packer's lane-interleaved order inside every work item (step 6b of pack.cpp): word 4j + i of a block of
    4 * LANES words holds the block's (i * nq + j)-th entry in ascending order, nq = quads in the block."""
    out = a.copy()
    for b, e in zip(item_off[:-1], item_off[1:]):
        B = 128 if e - b > item_len else 32
        for k in range(b, e, B):
            m = min(B, e - k)
            nq = m // 4
            r = np.arange(m)
            out[k + r] = a[k + 4 * (r % nq) + r // nq]
    return out


def bits(mask):
    return ((mask[:, None] >> np.arange(8)[None, :]) & 1).astype(np.float64)


def em_update_model4(arr, info, T, theta_T8, efflen_T8, unit=False):
    rowptr = arr["rowptr"].astype(np.int64)
    pairs = arr["pairs"].astype(np.int64)
    locus, mask = pairs & 0xFFFFFF, pairs >> 24
    cls_of_pair = np.repeat(np.arange(info["n_classes"]), np.diff(rowptr))
    x = (bits(mask) * (1.0 if unit else theta_T8[locus])).sum(axis=1)
    s = np.bincount(cls_of_pair, weights=x, minlength=info["n_classes"])
    w = np.append(arr["count"] / s, 0.0)  # trailing zero slot: padding entries point at it
    idx, emask = unpack_entries(arr["ent_cls"], info["entry_bytes"])
    item_off = arr["item_off"].astype(np.int64)
    item_of_entry = np.repeat(np.arange(info["n_items"]), np.diff(item_off))
    contrib = bits(emask) * w[idx][:, None]
    wit = np.zeros((info["n_items"], 8))
    np.add.at(wit, item_of_entry, contrib)
    lip = arr["locus_item_ptr"].astype(np.int64)
    locus_of_item = np.repeat(np.arange(T), np.diff(lip))
    W = np.zeros((T, 8))
    np.add.at(W, locus_of_item, wit)
    acc = W if unit else theta_T8 * W
    return acc, acc / efflen_T8
