"""Simulation clock and alarms, mirror of src/infra/TimeManager.jl (ESMF-like; drives ocn_run_loop).

  Clock, setCurrentTime!, changeTimeStep!, attachAlarm!, advance!      TimeManager.jl:5-63
  OneTimeAlarm, PeriodicAlarm, Alarm                                   TimeManager.jl:81-123
  isRinging, updateStatus!, rename!, stop!, reset!                     TimeManager.jl:126-171
  mpas_create_clock                                                    TimeManager.jl:173-189

`Period(value, unit)` stands for Julia's single-unit `Dates.Period`s (Year ... Second); `datetime + Period`
follows Julia's calendar arithmetic (months and years clamp the day to the end of the month).
"""
from __future__ import annotations

import calendar
import datetime as _dt

from ._lib import MokaError

_SECONDS = {"day": 86400, "hour": 3600, "minute": 60, "second": 1}


class Period:
    """Year(n) / Month(n) / Day(n) / Hour(n) / Minute(n) / Second(n)."""

    __slots__ = ("value", "unit")

    def __init__(self, value: int, unit: str):
        if unit not in ("year", "month", "day", "hour", "minute", "second"):
            raise MokaError(f"unknown period unit {unit}")
        self.value, self.unit = int(value), unit

    def __eq__(self, other):
        return isinstance(other, Period) and (self.value, self.unit) == (other.value, other.unit)

    def __hash__(self):
        return hash((self.value, self.unit))

    def __repr__(self):
        return f"{self.unit.capitalize()}({self.value})"

    def __radd__(self, t: _dt.datetime) -> _dt.datetime:
        if self.unit in _SECONDS:
            return t + _dt.timedelta(seconds=self.value * _SECONDS[self.unit])
        months = self.value * (12 if self.unit == "year" else 1)
        y, m0 = divmod(t.year * 12 + (t.month - 1) + months, 12)
        day = min(t.day, calendar.monthrange(y, m0 + 1)[1])
        return t.replace(year=y, month=m0 + 1, day=day)

    def __rsub__(self, t: _dt.datetime) -> _dt.datetime:
        return t + Period(-self.value, self.unit)

    def seconds(self) -> float:
        """Dates.value(Second(period)): only fixed-length periods convert (mpas_ocean.jl:36)."""
        if self.unit not in _SECONDS:
            raise MokaError(f"cannot convert {self!r} to seconds")
        return float(self.value * _SECONDS[self.unit])


def Year(n): return Period(n, "year")          # noqa: E704
def Month(n): return Period(n, "month")        # noqa: E704
def Day(n): return Period(n, "day")            # noqa: E704
def Hour(n): return Period(n, "hour")          # noqa: E704
def Minute(n): return Period(n, "minute")      # noqa: E704
def Second(n): return Period(n, "second")      # noqa: E704


class Clock:
    """TimeManager.jl:5-30."""

    def __init__(self, startTime: _dt.datetime, timeStep: Period):
        self.startTime = startTime
        self.currTime = startTime
        self.prevTime = None
        self.nextTime = startTime + timeStep
        self.timeStep = timeStep
        self.alarms: dict[str, "AbstractAlarm"] = {}

    def __repr__(self):
        return (f"Simulation Clock with {len(self.alarms)} Alarms attached\n  Start Time   : {self.startTime}\n"
                f"  Current Time : {self.currTime}\n  Previous Time: {self.prevTime}\n  Next Time    : {self.nextTime}\n"
                f"  Timestep     : {self.timeStep}")


def setCurrentTime(clock: Clock, inCurrTime: _dt.datetime) -> None:
    """TimeManager.jl:32-41 (a time before the start is only logged by the reference and leaves the clock unchanged)."""
    if inCurrTime < clock.startTime:
        return
    clock.currTime = inCurrTime
    clock.prevTime = inCurrTime - clock.timeStep
    clock.nextTime = inCurrTime + clock.timeStep


def changeTimeStep(clock: Clock, timestep: Period) -> None:
    """TimeManager.jl:43-48."""
    clock.timeStep = timestep
    clock.nextTime = clock.currTime + timestep


def attachAlarm(clock: Clock, alarm: "AbstractAlarm") -> None:
    """TimeManager.jl:50-53."""
    clock.alarms[alarm.name] = alarm


def advance(clock: Clock) -> None:
    """TimeManager.jl:55-63."""
    clock.prevTime = clock.currTime
    clock.currTime = clock.nextTime
    clock.nextTime = clock.currTime + clock.timeStep
    for alarm in clock.alarms.values():
        updateStatus(alarm, clock.currTime)


class AbstractAlarm:
    name: str
    ringing: bool
    stopped: bool


class OneTimeAlarm(AbstractAlarm):
    """TimeManager.jl:81-92."""

    def __init__(self, name: str, alarmTime: _dt.datetime):
        self.name, self.ringing, self.stopped, self.ringTime = name, False, False, alarmTime


class PeriodicAlarm(AbstractAlarm):
    """TimeManager.jl:95-117: first ring one interval after `intervalStart`."""

    def __init__(self, name: str, alarmInterval: Period, intervalStart: _dt.datetime):
        self.name, self.ringing, self.stopped = name, False, False
        self.ringTime = intervalStart + alarmInterval
        self.ringInterval = alarmInterval
        self.ringTimePrev = None


def Alarm(name: str, a, b=None) -> AbstractAlarm:
    """TimeManager.jl:120-123."""
    return OneTimeAlarm(name, a) if b is None else PeriodicAlarm(name, a, b)


def isRinging(alarm: AbstractAlarm) -> bool:
    return alarm.ringing                                              # TimeManager.jl:126-128


def updateStatus(alarm: AbstractAlarm, currentTime: _dt.datetime) -> None:
    if alarm.ringTime == currentTime:                                 # TimeManager.jl:130-132: equality, not >=
        alarm.ringing = True


def rename(alarm: AbstractAlarm, newName: str) -> None:
    alarm.name = newName                                              # TimeManager.jl:134-136


def stop(alarm: AbstractAlarm) -> None:
    alarm.ringing = False                                             # TimeManager.jl:138-140


def reset(alarm: AbstractAlarm, inTime: _dt.datetime | None = None) -> None:
    """TimeManager.jl:143-186."""
    stop(alarm)
    if isinstance(alarm, OneTimeAlarm):
        if inTime is None:
            alarm.stopped = True
        else:
            alarm.ringTime = inTime
        return
    if inTime is None:
        alarm.ringTimePrev = alarm.ringTime
        alarm.ringTime = alarm.ringTimePrev + alarm.ringInterval
    elif inTime >= alarm.ringTime:                                    # an earlier time is only logged by the reference
        while alarm.ringTime <= inTime:
            alarm.ringTimePrev = alarm.ringTime
            alarm.ringTime = alarm.ringTimePrev + alarm.ringInterval


def mpas_create_clock(timeStep: Period, startTime: _dt.datetime, stopTime=None, runDuration=None) -> Clock:
    """TimeManager.jl:173-189."""
    if runDuration is None and stopTime is None:
        raise MokaError(" neither stopTime nor runDuration are specified")
    return Clock(startTime, timeStep)
