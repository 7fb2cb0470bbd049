"""CPU: host-side control plane mirrors -- the reference's own infrastructure tests restated
(test/infra/test_Config.jl, test/infra/test_timeManager.jl) plus the NetCDF mesh/state round trip."""
import datetime as dt
import os

import numpy as np
import pytest

import moka_b200 as mb
from moka_b200 import io_netcdf
from moka_b200.config import ConfigAdd, ConfigGet, ConfigRead, ConfigSet
from moka_b200.time_manager import (Alarm, Clock, Day, Hour, Minute, Month, Second, Year, advance, attachAlarm, changeTimeStep,
                                    isRinging, reset, setCurrentTime, stop)
from conftest import hex_mesh

HERE = os.path.dirname(os.path.abspath(__file__))


def test_config_values_periods_and_datetimes():
    """test/infra/test_Config.jl:14-44."""
    config = ConfigRead(os.path.join(HERE, "data", "test.yaml"))
    hmix = ConfigGet(config.namelist, "hmix")
    intervals = ConfigGet(config.streams, "intervals")
    datetimes = ConfigGet(config.streams, "datetimes")
    assert ConfigGet(hmix, "hmix_String") == "Restart_timestamp"
    assert ConfigGet(hmix, "hmix_Float") == 1.234567890
    assert ConfigGet(hmix, "hmix_None") == "none"
    assert ConfigGet(hmix, "hmix_On") is True and ConfigGet(hmix, "hmix_Off") is False
    assert ConfigGet(hmix, "hmix_Exp") == 1.0e25
    assert ConfigGet(intervals, "yearly_interval") == Year(1)
    assert ConfigGet(intervals, "monthly_interval") == Month(2)
    assert ConfigGet(intervals, "daily_interval") == Day(3)
    assert ConfigGet(intervals, "hourly_interval") == Hour(4)
    assert ConfigGet(intervals, "minutes_interval") == Minute(5)
    assert ConfigGet(intervals, "seconds_interval") == Second(6)
    for key, want in (("NO_HMS", (0, 0, 0)), ("NO_MS", (2, 0, 0)), ("NO_S", (2, 3, 0)), ("NO_H", (0, 3, 4)),
                      ("NO_HM", (0, 0, 4)), ("NO_HS", (0, 3, 0)), ("ALL_HMS", (2, 3, 4))):
        assert ConfigGet(datetimes, key) == dt.datetime(1, 1, 1, *want)
    # ConfigAdd / ConfigSet error behaviour (Config.jl:60-86)
    ConfigAdd(hmix, "new_option", 3)
    with pytest.raises(mb.MokaError):
        ConfigAdd(hmix, "new_option", 4)
    ConfigSet(hmix, "new_option", 5)
    assert ConfigGet(hmix, "new_option") == 5
    with pytest.raises(mb.MokaError):
        ConfigSet(hmix, "missing", 1)
    with pytest.raises(mb.MokaError):
        ConfigRead(os.path.join(HERE, "data", "does_not_exist.yaml"))


def test_clock_and_alarms_two_year_integration():
    """test/infra/test_timeManager.jl: 2-year integration at a 20-minute step, every alarm checked when due."""
    time0 = dt.datetime(2000, 1, 1)
    clock = Clock(time0, Hour(1))
    assert clock.currTime == time0 and clock.timeStep == Hour(1)
    one_time = {"2020-03-01": dt.datetime(2020, 3, 1), "2019-08-24": dt.datetime(2019, 8, 24), "New Year 2020": dt.datetime(2020, 1, 1)}
    ota = {k: Alarm(k, v) for k, v in one_time.items()}
    per = {"20min": Alarm("Every 20 minutes", Minute(20), time0), "1h": Alarm("Every hour", Hour(1), time0),
           "6h": Alarm("Every 6 hours", Hour(6), time0), "day": Alarm("Every day", Day(1), time0),
           "month": Alarm("Every month", Month(1), time0), "year": Alarm("Every year", Year(1), time0)}
    for a in list(ota.values()) + list(per.values()):
        attachAlarm(clock, a)
    changeTimeStep(clock, Minute(20))
    assert clock.timeStep == Minute(20)
    cur = dt.datetime(2019, 1, 1)
    setCurrentTime(clock, cur)
    assert clock.currTime == cur and clock.prevTime == dt.datetime(2018, 12, 31, 23, 40) and clock.nextTime == dt.datetime(2019, 1, 1, 0, 20)
    for a in per.values():
        reset(a, cur)
    stop_time = dt.datetime(2021, 1, 1)
    rang = {k: 0 for k in list(ota) + list(per)}
    while clock.currTime <= stop_time:
        advance(clock)
        t = clock.currTime
        for k, when in one_time.items():
            if t == when:
                assert isRinging(ota[k])
                stop(ota[k])
                rang[k] += 1
        due = {"20min": t.minute % 20 == 0 and t.second == 0, "1h": t.minute == 0 and t.second == 0,
               "6h": t.hour % 6 == 0 and t.minute == 0 and t.second == 0, "day": t.hour == 0 and t.minute == 0 and t.second == 0,
               "month": t.day == 1 and t.hour == 0 and t.minute == 0 and t.second == 0,
               "year": t.month == 1 and t.day == 1 and t.hour == 0 and t.minute == 0 and t.second == 0}
        for k, d in due.items():
            if d:
                assert isRinging(per[k]), (k, t)
                reset(per[k])                                        # so the next due time is a fresh ring, not a stale flag
                rang[k] += 1
            else:
                assert not isRinging(per[k]), (k, t)
    assert all(rang[k] == 1 for k in one_time)
    assert rang["year"] == 2 and rang["month"] == 24 and rang["day"] == 731
    # calendar arithmetic of the periods (Julia clamps the day to the end of the month)
    assert dt.datetime(2020, 1, 31) + Month(1) == dt.datetime(2020, 2, 29)
    assert dt.datetime(2020, 2, 29) + Year(1) == dt.datetime(2021, 2, 28)
    assert dt.datetime(2020, 3, 1) - Month(1) == dt.datetime(2020, 2, 1)


def test_mesh_file_roundtrip(tmp_path):
    """write_mesh_netcdf -> ReadHorzMesh returns every array unchanged; the layout on disk is the reference's."""
    from scipy.io import netcdf_file
    m = hex_mesh(12)
    ssh, u, h = mb.inertialGravityWave(m).initial_state()
    path = str(tmp_path / "mesh.nc")
    mb.write_mesh_netcdf(path, m, (ssh, u, h))
    f = mb.ReadHorzMesh(path)
    for k in io_netcdf._MESH_VARS:
        if k in m:
            assert np.array_equal(f[k], m[k]), k
    assert f["is_periodic"] == "YES" and f["nVertLevels"] == 1 and np.array_equal(f["restingThickness"], m["restingThickness"])
    s2, u2, h2 = io_netcdf.read_initial_state(path, m["nCells"], m["nEdges"])
    assert np.array_equal(s2, ssh) and np.array_equal(u2, u) and np.array_equal(h2, h)
    with netcdf_file(path, "r", mmap=False) as ds:
        assert ds.variables["edgesOnCell"].dimensions == ("nCells", "maxEdges")         # Julia reads (maxEdges, nCells)
        assert ds.variables["restingThickness"].dimensions == ("Time", "nCells", "nVertLevels")   # [:,:,1], VertMesh.jl:57
        assert ds.variables["normalVelocity"].dimensions == ("Time", "nEdges", "nVertLevels")     # PrognosticVars.jl:98
    # a mesh that is not flagged periodic and carries no boundary mask is rejected like VertMesh.jl:50-52
    bad = dict(f)
    bad["is_periodic"] = "NO"
    bad.pop("boundaryEdge", None)
    with pytest.raises(mb.MokaError, match="non-periodic"):
        mb.VerticalMesh(path, bad, backend=None)
