"""
ORACLE — TEST INFRASTRUCTURE ONLY.

The north-star parity rule for free-running chains (BASELINE.json: "decoded token ids after rounding must be bit-exact,
except where a stated top-2 distance margin is below tolerance"), applied to EVERY rounding call of a chain.

A chain is a sequence of rounding calls (rounding.py:31-47 invoked at diffusion.py:322); the reference side supplies the
ids it chose and the squared-distance gap between its nearest and second-nearest embedding at every call and position.
A position (b, l) may disagree with the reference only if, at the FIRST call where its id differs,
  (a) the reference's own top-2 margin there was below `tol` (the bf16 denoiser legitimately flips a near-tie), or
  (b) an earlier, rule-(a) flip in the SAME sequence had already changed x_t (every later model output of that sequence
      then differs through attention; such positions are reported as `downstream` and bounded by `tol_downstream`).
Anything else is a violation.  `tol` may be one number or one number per call: for an epsilon-predicting model the rounded
quantity is x0 = sqrt_recip[t] x_t - sqrt_recipm1[t] eps (diffusion.py:194-199), so the denoiser's bf16 error reaches the
distances multiplied by sqrt_recipm1[t] and the tolerance of call t scales with it.  Prefix positions (mask == 0) are excluded: their rounding result never reaches x_{t-1}
(diffusion.py:394-397 re-applies x_start).
"""
import numpy as np


def chain_report(ref_ids, ref_margin, got_ids, free, tol, tol_downstream=None):
    """ref_ids / got_ids: int [S, B, L]; ref_margin: float [S, B, L]; free: bool [B, L] (mask != 0).
    Returns a dict of counts; `violations` must be 0 for parity."""
    ref_ids = np.asarray(ref_ids).astype(np.int64)
    got_ids = np.asarray(got_ids).astype(np.int64)
    ref_margin = np.asarray(ref_margin, dtype=np.float64)
    S, B, L = ref_ids.shape
    assert got_ids.shape == ref_ids.shape == ref_margin.shape and free.shape == (B, L)
    differ = (ref_ids != got_ids) & free[None]
    any_div = differ.any(axis=0)                                   # [B, L]
    first = np.where(any_div, differ.argmax(axis=0), S)            # first divergent call per position (S = never)
    row_first = first.min(axis=1)                                  # first divergent call of each sequence
    bi, li = np.nonzero(any_div)
    tol = np.broadcast_to(np.asarray(tol, dtype=np.float64), (S,))
    td = tol if tol_downstream is None else np.broadcast_to(np.asarray(tol_downstream, dtype=np.float64), (S,))
    m_first = ref_margin[first[bi, li], bi, li] / tol[first[bi, li]] * tol.min()      # in units of the smallest tolerance
    primary = first[bi, li] == row_first[bi]                       # diverged while the sequence was still identical
    viol_primary = primary & (m_first >= tol.min())
    viol_down = (~primary) & (m_first >= tol.min() * (td / tol)[first[bi, li]])
    n_free = int(free.sum())
    return {
        "calls": S, "free_positions": n_free,
        "id_agreement_all_calls": float(1.0 - differ.sum() / max(1, S * n_free)),
        "positions_ever_divergent": int(any_div.sum()),
        "primary": int(primary.sum()), "downstream": int((~primary).sum()),
        "max_margin_primary": float(m_first[primary].max()) if primary.any() else 0.0,
        "max_margin_downstream": float(m_first[~primary].max()) if (~primary).any() else 0.0,
        "violations": int(viol_primary.sum() + viol_down.sum()),
        "sequences_clean": int((row_first == S).sum()), "sequences": B,
        "never_divergent": ~any_div & free,
    }


def token_report(ref_tokens, got_tokens, free, never_divergent, ref_logit_margin=None, logit_tol=1e-3):
    """Decoded tokens (run/sample.py:219-220).  Positions whose rounding ids agreed at EVERY call carry the same x_0 up to
    fp32 rounding, so their tokens must be identical (unless the reference's own top-2 logit gap is below `logit_tol`)."""
    ref_tokens, got_tokens = np.asarray(ref_tokens), np.asarray(got_tokens)
    same = ref_tokens == got_tokens
    must = never_divergent.copy()
    if ref_logit_margin is not None:
        must &= np.asarray(ref_logit_margin) >= logit_tol
    return {"token_agreement": float(same[free].mean()) if free.any() else 1.0,
            "token_agreement_all": float(same.mean()),
            "token_violations": int((~same & must).sum())}


def oracle_chain_trace(record, E):
    """(ids [S, B, L], margin [S, B, L]) of an oracle run from its `record` list (musediff_oracle.p_sample_loop /
    ddim_sample_loop with record=[]), recomputed with the oracle's own rounding."""
    import musediff_oracle as O
    ids, margins = [], []
    for r in record:
        mo = r["model_output"]
        idx, dist = O.efficient_knn(E, mo)
        ids.append(idx.reshape(mo.shape[:-1]))
        margins.append(O.top2_margin(dist).reshape(mo.shape[:-1]))
    return np.stack(ids), np.stack(margins)


def trace_to_numpy(trace, B, L):
    """diffusion.rounding_trace -> (ids [S, B, L], margin [S, B, L])."""
    ids = np.stack([t[0].view(B, L).cpu().numpy() for t in trace])
    mar = np.stack([t[1].view(B, L).cpu().numpy() for t in trace])
    return ids, mar
