"""The `--correspond` wire format (src/main.rs:182-244): what the web client's POST /write/ receives.

correspondence(engine, reminder, bound) mirrors `main::correspondence`: reconstruct the world from its
runes, search (the kickoff family on the device), and answer with a `Postcard` or, when the game is over,
a `LastMissive` -- JSON laid out exactly as rustc-serialize writes those structs (field order as declared,
no whitespace, unit enums as strings, `None` as null; `TransitPatch` / `RelaxedLocale`, src/life.rs:22-37,
src/space.rs:6-18).
"""
import json
import time

from . import runes

TEAMS = ("Orange", "Blue")
JOBS = ("Servant", "Pony", "Scholar", "Cop", "Princess", "Figurehead")


class Depth:  # LookaheadBound::Depth(depth, extension), main.rs:90-94
    def __init__(self, depth, extension=None):
        self.depth, self.extension = depth, extension


class DepthSequence:  # LookaheadBound::DepthSequence
    def __init__(self, depths):
        self.depths = list(depths)

    @classmethod
    def from_depiction(cls, depiction):  # main.rs:108-117
        return cls(int(d) for d in depiction.split(","))


class Seconds:  # LookaheadBound::Seconds
    def __init__(self, seconds):
        self.seconds = seconds


def _agent(byte):
    return {"team": TEAMS[byte >> 3], "job_description": JOBS[byte & 7]}


def _relaxed(loc):
    return {"rank": int(loc) >> 4, "file": int(loc) & 15}


def transit_patch(team, job, whence, whither):
    """TransitPatch::from(Patch), life.rs:22-37"""
    return {"star": {"team": TEAMS[team], "job_description": JOBS[job]}, "whence": _relaxed(whence), "whither": _relaxed(whither)}


def abbreviated_pagan_movement_rune(job, whence, whither):
    """life.rs:56-70 -- O–O / O–O–O use en dashes, like the reference"""
    if job == 5 and abs((int(whence) & 15) - (int(whither) & 15)) == 2:
        return {6: "O–O", 2: "O–O–O"}[int(whither) & 15]
    return ("" if job == 0 else "PNBRQK"[job]) + "abcdefgh"[int(whither) & 15] + str((int(whither) >> 4) + 1)


def forecast(engine, world, bound):
    """main.rs:152-180 -> (forecasts, depth, thinking time in ms)"""
    start = time.perf_counter()
    if isinstance(bound, Depth):
        forecasts, depth = engine.deep_forecast(world, bound.depth, quiet=bound.extension), bound.depth
    elif isinstance(bound, DepthSequence):
        forecasts, depth = engine.fixed_depth_sequence_kickoff(world, bound.depths), bound.depths[-1]
    else:
        forecasts, depth = engine.iterative_deepening_kickoff(world, bound.seconds)
    return forecasts, depth, int((time.perf_counter() - start) * 1e3)


def _encode(obj):
    return json.dumps(obj, separators=(",", ":"), ensure_ascii=False)


def correspondence(engine, reminder, bound):
    """main.rs:199-244"""
    in_medias_res = runes.reconstruct(reminder)
    forecasts, depth, sidereal = forecast(engine, in_medias_res, bound)
    if len(forecasts):
        determination = forecasts[0]["commit"]
        _, replies = engine.lookahead(determination["tree"])
        if len(replies) == 0:
            endangered = bool(engine.in_critical_endangerment([determination["tree"]], 0)[0])
            return _encode({"the_triumphant": "Blue" if endangered else None})
        hosp = int(determination["hospitalization"])
        return _encode({
            "world": runes.preserve(determination["tree"]),
            "patch": transit_patch(int(determination["team"]), int(determination["job"]), determination["whence"], determination["whither"]),
            "hospitalization": None if hosp == 0xFF else _agent(hosp),
            "thinking_time": sidereal,
            "depth": depth,
            "counterreplies": [transit_patch(int(c["team"]), int(c["job"]), c["whence"], c["whither"]) for c in replies],
            "rosetta_stone": abbreviated_pagan_movement_rune(int(determination["job"]), determination["whence"], determination["whither"]),
        })
    endangered = bool(engine.in_critical_endangerment([in_medias_res], 1)[0])
    return _encode({"the_triumphant": "Orange" if endangered else None})
