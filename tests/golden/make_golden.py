"""Writes tests/golden/nexmark_golden.json: expected NEXMark q1-q8 outputs (q6: oracle only so far) on a small seeded input.

The vectors are produced by the CPU oracle (oracle/) AFTER it has been cross-checked, in this script,
against the independent Arrow C++ implementation (oracle/acero_ref.py); the reference itself cannot run
here (no Rust toolchain), which is why the header of oracle/__init__.py says "parity unpinned".
Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import oracle                      # noqa: E402
from oracle import acero_ref       # noqa: E402
from flock_b200 import nexgen, plans  # noqa: E402

N_EVENTS, SEED, BATCH_ROWS = 50_000, 1234, 4096


def main():
    ev = nexgen.generate(N_EVENTS, seed=SEED, batch_rows=BATCH_ROWS)
    out = {"n_events": N_EVENTS, "seed": SEED, "batch_rows": BATCH_ROWS, "queries": {}}
    for q in ["q1", "q2", "q3", "q4", "q5", "q6", "q7", "q8"]:
        sources = [[ev[r]] for r in plans.SOURCES[q]]
        got = oracle.execute_plan(plans.QUERIES[q](), sources)
        rels = [ev[r] for r in dict.fromkeys(plans.SOURCES[q])]
        oracle.assert_tables_equal(got, acero_ref.QUERIES[q](*rels))      # two implementations agree
        t = oracle.canonical(got)
        rows = json.dumps(t.to_pylist(), default=str, sort_keys=True)
        out["queries"][q] = {"num_rows": t.num_rows, "columns": t.schema.names,
                             "head": json.loads(json.dumps(t.slice(0, 5).to_pylist(), default=str)),
                             "digest": hashlib.sha256(rows.encode()).hexdigest()}
    path = Path(__file__).parent / "nexmark_golden.json"
    path.write_text(json.dumps(out, indent=1) + "\n")
    print("wrote", path, {q: v["num_rows"] for q, v in out["queries"].items()})


if __name__ == "__main__":
    main()
