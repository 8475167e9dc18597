"""Second, independent CPU implementation of the NEXMark queries on Arrow C++ compute / Acero (pyarrow).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  It exists to pin the C++ restatement: two
implementations written against different libraries must agree bit-exactly (after the canonical sort)
before either is trusted as the oracle.  It works from the SQL text
(benchmarks/src/nexmark/query/qN.sql), not from the physical plan JSON, so it also cross-checks the
hand-written plans in flock_b200/plans.py.  Label: "Arrow C++ kernels -- not DataFusion".
"""
from __future__ import annotations

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc


def _table(batches) -> pa.Table:
    return pa.Table.from_batches(list(batches)).combine_chunks()


def q1(bid) -> pa.Table:
    """SELECT auction, bidder, 0.908 * price AS price, b_date_time FROM bid"""
    t = _table(bid)
    price = pc.multiply(pa.scalar(0.908, pa.float64()), pc.cast(t["price"], pa.float64()))
    return pa.table({"auction": t["auction"], "bidder": t["bidder"], "price": price, "b_date_time": t["b_date_time"]})


def q2(bid) -> pa.Table:
    """SELECT auction, price FROM bid WHERE auction % 123 = 0  (truncated remainder, like Rust's %)"""
    t = _table(bid)
    a = t["auction"].to_numpy().astype(np.int64)
    mask = np.fmod(a, 123) == 0
    return t.select(["auction", "price"]).filter(pa.array(mask))


def q3(auction, person) -> pa.Table:
    """SELECT name, city, state, a_id FROM auction JOIN person ON seller = p_id
       WHERE category = 10 AND (state = 'or' OR state = 'id' OR state = 'ca')"""
    a = _table(auction).select(["a_id", "seller", "category"])
    p = _table(person).select(["p_id", "name", "city", "state"])
    a = a.filter(pc.equal(pc.cast(a["category"], pa.int64()), 10))
    p = p.filter(pc.is_in(p["state"], value_set=pa.array(["or", "id", "ca"])))
    j = a.join(p, keys="seller", right_keys="p_id", join_type="inner")
    return j.select(["name", "city", "state", "a_id"])


def q5(bid) -> pa.Table:
    """AuctionBids(auction, COUNT(*) num) JOIN MaxBids(MAX(num) maxn) ON num = maxn"""
    t = _table(bid)
    counts = t.group_by("auction").aggregate([([], "count_all")])
    num = pc.cast(counts["count_all"], pa.uint64())
    mx = pc.max(num)
    out = pa.table({"auction": counts["auction"], "num": num})
    return out.filter(pc.equal(num, mx))


def q8(person, auction) -> pa.Table:
    """P(SELECT p_id, name GROUP BY p_id, name) JOIN A(SELECT seller GROUP BY seller) ON p_id = seller"""
    p = _table(person).select(["p_id", "name"]).group_by(["p_id", "name"]).aggregate([])
    sellers = pc.unique(_table(auction)["seller"].combine_chunks())
    return p.filter(pc.is_in(p["p_id"], value_set=sellers)).select(["p_id", "name"])


def q4(auction, bid) -> pa.Table:
    """SELECT category, AVG(final) FROM (SELECT MAX(price) final, category FROM auction JOIN bid ON a_id = auction
       WHERE b_date_time BETWEEN a_date_time AND expires GROUP BY a_id, category) GROUP BY category  -- numpy, from the SQL"""
    a = _table(auction)
    b = _table(bid)
    a_id = a["a_id"].to_numpy()
    order = np.argsort(a_id, kind="stable")
    keys = a_id[order]
    auc = b["auction"].to_numpy()
    lo = np.searchsorted(keys, auc, "left")
    hi = np.searchsorted(keys, auc, "right")
    reps = hi - lo                                                     # matches per bid (a_id is unique in NEXMark, but stay general)
    bi = np.repeat(np.arange(len(auc)), reps)
    ai = order[np.concatenate([np.arange(l, h) for l, h in zip(lo[reps > 0], hi[reps > 0])])] if reps.sum() else np.zeros(0, np.int64)
    ts = b["b_date_time"].cast(pa.int64()).to_numpy()[bi]
    keep = (ts >= a["a_date_time"].cast(pa.int64()).to_numpy()[ai]) & (ts <= a["expires"].cast(pa.int64()).to_numpy()[ai])
    ai, bi = ai[keep], bi[keep]
    inner = pa.table({"a_id": a_id[ai], "category": a["category"].to_numpy()[ai], "price": b["price"].to_numpy()[bi]})
    inner = inner.group_by(["a_id", "category"]).aggregate([("price", "max")])
    cat = inner["category"].to_numpy()
    final = inner["price_max"].to_numpy().astype(np.float64)
    cats = np.unique(cat)
    avg = np.array([final[cat == c].sum() / np.float64((cat == c).sum()) for c in cats], np.float64)   # exact integer sums < 2^53
    return pa.table({"category": pa.array(cats, pa.int32()), "AVG(Q.final)": pa.array(avg, pa.float64())})


def q7(bid) -> pa.Table:
    """SELECT auction, price, bidder, b_date_time FROM bid JOIN (SELECT MAX(price) maxprice FROM bid) ON price = maxprice"""
    t = _table(bid)
    mx = pc.max(t["price"])
    return t.filter(pc.equal(t["price"], mx)).select(["auction", "price", "bidder", "b_date_time"])


def q6(auction, bid) -> pa.Table:
    """SELECT seller, AVG(price) FROM (last ten winning bids per seller, by closing-bid time) -- q6.sql, from the SQL text.
    Ties (the reference leaves them undefined) are broken like oracle.sort_batch: by the remaining columns, ascending --
    i.e. among equal maximum prices of an auction the earliest bid wins, among equal bid times of a seller the lower price
    ranks first."""
    a = _table(auction)
    b = _table(bid)
    a_id = a["a_id"].to_numpy()
    order = np.argsort(a_id, kind="stable")
    keys = a_id[order]
    auc = b["auction"].to_numpy()
    pos = np.searchsorted(keys, auc, "left")
    hit = (pos < len(keys)) & (keys[np.minimum(pos, len(keys) - 1)] == auc)      # a_id is unique in NEXMark
    bi = np.nonzero(hit)[0]
    ai = order[pos[hit]]
    ts = b["b_date_time"].cast(pa.int64()).to_numpy()[bi]
    keep = (ts >= a["a_date_time"].cast(pa.int64()).to_numpy()[ai]) & (ts <= a["expires"].cast(pa.int64()).to_numpy()[ai])
    ai, bi, ts = ai[keep], bi[keep], ts[keep]
    price = b["price"].to_numpy()[bi].astype(np.int64)
    aid = a_id[ai]
    o = np.lexsort((ts, -price, aid))                                           # a_id, price DESC, then earliest bid
    first = np.ones(len(o), bool)
    first[1:] = aid[o][1:] != aid[o][:-1]
    w = o[first]                                                                # the winning bid of every auction
    seller = a["seller"].to_numpy()[ai][w]
    wprice, wts = price[w], ts[w]
    o2 = np.lexsort((wprice, -wts, seller))                                     # seller, b_date_time DESC, then lower price
    s2 = seller[o2]
    start = np.ones(len(o2), bool)
    start[1:] = s2[1:] != s2[:-1]
    idx = np.arange(len(o2))
    rank = idx - np.maximum.accumulate(np.where(start, idx, 0)) + 1
    last10 = rank <= 10
    s3, p3 = s2[last10], wprice[o2][last10]
    sellers, inv = np.unique(s3, return_inverse=True)
    sums = np.bincount(inv, weights=p3.astype(np.float64))
    cnts = np.bincount(inv)
    return pa.table({"seller": pa.array(sellers, pa.int32()), "AVG(R.price)": pa.array(sums / cnts.astype(np.float64), pa.float64())})


QUERIES = {"q1": q1, "q2": q2, "q3": q3, "q4": q4, "q5": q5, "q6": q6, "q7": q7, "q8": q8}
