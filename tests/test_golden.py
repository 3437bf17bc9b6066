"""Oracle outputs on the 3 000-cell seeded planet against the committed checksums (tests/golden/)."""
import json
import os


def test_oracle_matches_committed_checksums():
    from tests.golden.make_golden import compute
    want = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_checksums.json")))
    got = compute()
    assert set(got) == set(want)
    bad = [k for k in want if got[k] != want[k]]
    assert not bad, f"oracle output changed for: {bad}"
