"""CPU: host-side pieces that need no GPU — validation of day traces at the packing boundary (what the reference lets
acnsim raise at run time is refused here at construction), the lazily built infos sequence of the vector adapters, and
the env-id sharding of the multi-GPU path."""
import numpy as np
import pytest

from sustaingym_b200.data.ev_traces import pack_sessions
from sustaingym_b200.envs.vector import LazyInfos


def test_pack_sessions_refuses_what_acnsim_would_raise_on():
    ok = [(3, 40, 38, 12.5, 7), (40, 60, 61, 3.0, 7), (5, 9, 9, 200.0, 0)]       # chained on station 7; over the cap
    t = pack_sessions([ok, []], 54)
    assert t.sessions_of_day(1) == [] and len(t.sessions_of_day(0)) == 3
    assert max(s[3] for s in t.sessions_of_day(0)) == 100.0                        # requested energy capped (event_generation.py:169-170)
    with pytest.raises(ValueError, match="station index"):
        pack_sessions([[(1, 2, 2, 1.0, 54)]], 54)
    with pytest.raises(ValueError, match="arrival <= departure"):
        pack_sessions([[(10, 5, 6, 1.0, 0)]], 54)
    with pytest.raises(ValueError, match="arrival <= departure"):
        pack_sessions([[(10, 288, 287, 1.0, 0)]], 54)
    with pytest.raises(ValueError, match="overlap"):                               # StationOccupiedError in acnsim
        pack_sessions([[(10, 50, 50, 1.0, 3), (49, 80, 80, 1.0, 3)]], 54)
    with pytest.raises(ValueError, match="max_sessions"):
        pack_sessions([[(1, 2, 2, 1.0, 0), (3, 4, 4, 1.0, 0), (5, 6, 6, 1.0, 0)]], 54, max_sessions=2)


def test_lazy_infos_builds_dicts_only_for_finished_envs_and_only_on_access():
    infos = LazyInfos(6)
    assert len(infos) == 6 and all(len(x) == 0 for x in infos)
    fetched = []

    def fetch():
        fetched.append(1)
        return np.arange(2 * 4, dtype=np.float32).reshape(2, 4) + 100          # rows of the two finished envs

    infos._set(np.array([1, 4]), fetch, lambda row: {"x": row.copy()})
    assert not fetched                                                         # nothing fetched yet
    assert infos[0] == {} and infos[5] == {} and not fetched                   # unfinished envs never trigger the fetch
    d = infos[4]
    assert fetched == [1] and d["TimeLimit.truncated"] is False and np.array_equal(d["terminal_observation"]["x"], [104, 105, 106, 107])
    assert infos[4] is d and infos[-2] is d                                    # cached; negative indices
    assert np.array_equal(infos[1]["terminal_observation"]["x"], [100, 101, 102, 103]) and fetched == [1]
    assert [len(x) for x in infos[0:6]] == [0, 2, 0, 0, 2, 0]
    infos._clear()
    assert all(len(x) == 0 for x in infos)


def test_shard_range_partitions_contiguously():
    from sustaingym_b200.dist import shard_range
    for total, world in ((65536, 8), (262144, 8), (32768, 3), (10, 4), (3, 8)):
        parts = [shard_range(total, r, world) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [hi - lo for lo, hi in parts]
        assert max(sizes) - min(sizes) <= 1
