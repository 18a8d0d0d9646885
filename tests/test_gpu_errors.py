"""Error behaviour of the C ABI on a GPU: invalid input and exhausted capacity are reported as return codes with
a message (raised as PocError by the Python layer), never as crashes or silent corruption."""
import numpy as np
import pytest

import util as U

pytestmark = pytest.mark.gpu


def _engine(P, st, **kw):
    n, stat, crim = P.CrimeEngine.capacity_for([st], kw.pop("headroom", 1000))
    return P.CrimeEngine(1, n, stat, crim)


def test_invalid_inputs_are_rejected():
    import pyproton_oc_b200 as P
    st = U.micro_net_state()
    # unsorted row
    bad = dict(st)
    bad["col_5"] = st["col_5"].copy()
    rp = st["rowptr_5"]
    a = int(np.flatnonzero(np.diff(rp) >= 2)[0]) if (np.diff(rp) >= 2).any() else None
    if a is not None:
        bad["col_5"][rp[a]:rp[a] + 2] = bad["col_5"][rp[a]:rp[a] + 2][::-1]
        with _engine(P, st) as E:
            E.set_params(P.make_params({}))
            with pytest.raises(P.PocError, match="invalid input"):
                E.upload_state(0, bad)
    # neighbour slot out of range
    bad = dict(st)
    bad["col_7"] = st["col_7"].copy()
    bad["col_7"][0] = 99
    with _engine(P, st) as E:
        E.set_params(P.make_params({}))
        with pytest.raises(P.PocError, match="invalid input"):
            E.upload_state(0, bad)
    # attribute out of the packed range
    bad = dict(st)
    bad["job_level"] = st["job_level"].copy()
    bad["job_level"][3] = 40
    with _engine(P, st) as E:
        E.set_params(P.make_params({}))
        with pytest.raises(P.PocError, match="invalid input"):
            E.upload_state(0, bad)
    # more agents than the ctx was created for
    with P.CrimeEngine(1, 10, 100, 100) as E:
        with pytest.raises(P.PocError):
            E.upload_state(0, st)
    # overlapping age cells / radius out of range
    with _engine(P, st) as E:
        with pytest.raises(P.PocError, match="overlap"):
            E.set_params(P.make_params({}, c_range=[(1, 0, 20, 0.1), (1, 15, 30, 0.2)]))
        with pytest.raises(P.PocError):
            E.set_params(P.make_params({"max_accomplice_radius": 99}))
    # stage hooks validate slots
    with _engine(P, st) as E:
        E.set_params(P.make_params({}))
        E.upload_state(0, st)
        with pytest.raises(P.PocError):
            E.rings([500], 2, True)
        with pytest.raises(P.PocError):
            E.embeddedness([-1])
        # and a good call still works afterwards (argument errors are not sticky)
        assert E.rings([0], 2, True)[0].tolist() == [6]


def test_criminal_capacity_overflow_is_reported():
    import pyproton_oc_b200 as P
    fx = U.load_fixture(U.GOLDEN + "/tick_n1000_s11_oc200_t3.npz")
    st = U.pre_state(fx)
    n, stat, crim = P.CrimeEngine.capacity_for([st], 0)   # no headroom for new co-offender links
    with P.CrimeEngine(1, n, stat, crim) as E:
        E.set_params(P.make_params(U.fixture_params(fx)))
        E.upload_state(0, st)
        E.set_crime_multiplier(float(fx["crime_multiplier"]))
        E.run_ticks(int(fx["tick"]), 30, [1])
        cn, _ = E.counters()
        assert cn[0, 9] == 13   # CN_ERROR: criminal layer capacity
        # the state is still readable
        after = E.download_state(0)
        assert len(after["co_off"]) <= crim


def test_queued_uploads_report_errors_at_sync_and_result_buffers_are_checked():
    """upload_state(wait=False) returns before the device-side validation ran: the error surfaces at sync() and the
    ctx stays usable.  Caller-provided result buffers are validated on the host."""
    import pyproton_oc_b200 as P
    st = U.micro_net_state()
    bad = dict(st)
    bad["col_7"] = st["col_7"].copy()
    bad["col_7"][0] = 99
    with _engine(P, st) as E:
        E.set_params(P.make_params({}))
        E.upload_state(0, bad, wait=False)
        with pytest.raises(P.PocError, match="invalid input"):
            E.sync()
        E.upload_state(0, st, wait=False)
        E.sync()
        got = E.download_agents(0, out=E.agent_buffers(len(st["alive"])))
        assert np.array_equal(got["age"], st["age"])
        with pytest.raises(ValueError):
            E.download_agents(0, out={k: v[:3] for k, v in E.agent_buffers(len(st["alive"])).items()})
        with pytest.raises(ValueError):
            E.run_ticks(1, 2, [1], out=np.zeros((1, 3, len(P.REPORTER_NAMES))))
        rep = np.zeros((1, 2, len(P.REPORTER_NAMES)))
        assert E.run_ticks(1, 2, [1], out=rep) is rep and rep[0, 1, 0] >= rep[0, 0, 0]


def test_failed_ctx_creation_releases_everything_and_keeps_its_message():
    import pyproton_oc_b200 as P
    st = U.micro_net_state()
    with pytest.raises(P.PocError, match="shared-memory"):
        P.CrimeEngine(1, 2_000_000, 10, 10)          # fails half-way through poc_ctx_create
    with _engine(P, st) as E:                          # the next ctx of this process is unaffected
        E.set_params(P.make_params({}))
        E.upload_state(0, st)
        assert E.rings([0], 2, True)[0].tolist() == [6]
