"""The console game of the reference binary (src/main.rs:346-456) over the B200 leaf layer.

Every turn the world is depicted (life.rs:1088-1128), the choices are listed -- the legal commits when no lookahead
bound was given, the forecasts (score and representative variation) when one was -- and a choice is read:
an index, or `quit`.  Nothing here computes a move: commits come from `Engine.lookahead`, forecasts from
`correspondence.forecast`, the "+" of a movement rune from `Engine.in_critical_endangerment`.
"""
import sys

from . import correspondence as co
from . import runes

FIGURINES = "♙♘♗♖♕♔♟♞♝♜♛♚"  # identity.rs:162-177, plane order
_PREFIX = ("", "N", "B", "R", "Q", "K")  # to_pagan_movement_rune_prefix, identity.rs:149-160
_SCENERY = (58, 17)  # life.rs:1090
_PAINT = ("\x1b[1;31m", "\x1b[36m")  # Team::figurine_paintjob, identity.rs:27-32: Orange bold red, Blue cyan
_RESET = "\x1b[0m"


def algebraic(locale):
    return "abcdefgh"[int(locale) & 15] + str((int(locale) >> 4) + 1)


def _figurine(agent, colour):
    team, job = int(agent) >> 3, int(agent) & 7
    rune = FIGURINES[6 * team + job]
    return _PAINT[team] + rune + _RESET if colour else rune


def depict_world(world, colour=True):
    """`impl fmt::Display for WorldState` (life.rs:1088-1128): ranks 8 down to 1, two columns a locale, solid runes."""
    lines = ["    a b c d e f g h"]
    scene = 0
    for rank in range(7, -1, -1):
        row = " %d " % (rank + 1)
        for file in range(8):
            bit = 1 << (8 * rank + file)
            cell = "  "
            for plane in range(12):
                if int(world["plane"][plane]) & bit:
                    cell = FIGURINES[6 + plane % 6] + " "  # to_solid_display_rune: Blue's runes for both teams
                    if colour:
                        cell = _PAINT[plane // 6] + cell
                    break
            row += ("\x1b[48;5;%dm" % _SCENERY[scene] + cell + _RESET) if colour else cell
            scene ^= 1
        scene ^= 1
        lines.append(row)
    return "\n".join(lines) + "\n"


def pagan_movement_rune(commit, endangering):
    """Commit::pagan_movement_rune (life.rs:85-108)"""
    job, whence, whither = int(commit["job"]), commit["whence"], commit["whither"]
    if int(commit["hospitalization"]) != 0xFF:
        movement = _PREFIX[job] + "x" + algebraic(whither)
    else:
        movement = co.abbreviated_pagan_movement_rune(job, whence, whither)
    ascension = "" if int(commit["ascension"]) == 0xFF else "=" + _PREFIX[int(commit["ascension"]) & 7]
    return movement + ("+" if endangering else "") + ascension


def depict_commit(commit, endangering, colour=True):
    """`impl fmt::Display for Commit` (life.rs:111-155)"""
    report = ""
    if int(commit["hospitalization"]) != 0xFF:
        report += ", stunning " + _figurine(commit["hospitalization"], colour)
    if int(commit["ascension"]) != 0xFF:
        form = int(commit["ascension"])
        verb = {1: "transforming into", 3: "being brevetted to", 2: "transitioning into", 4: "transitioning into"}[form & 7]
        report += ", %s %s" % (verb, _figurine(form, colour))
    star = (int(commit["team"]) << 3) | int(commit["job"])
    return "%s%s (%s from %s to %s%s)" % ("" if int(commit["team"]) == 0 else "..", pagan_movement_rune(commit, endangering),
                                         _figurine(star, colour), algebraic(commit["whence"]), algebraic(commit["whither"]), report)


def variation_format(forecast):
    """pagan_variation_format (mind.rs:175-180)"""
    n = int(forecast["variation_len"])
    return " ".join(co.abbreviated_pagan_movement_rune(int(p["job"]), p["whence"], p["whither"]) for p in forecast["variation"][:n])


def _endangering(engine, commits):
    """which commits leave the opposition's figurehead in critical endangerment (the "+" of the rune)"""
    if not len(commits):
        return []
    trees = [c["tree"] for c in commits]
    return list(engine.in_critical_endangerment(trees, [int(t["initiative"]) for t in trees]))


def play(engine, world, bound=None, stdin=None, stdout=None, colour=True):
    """The loop of main.rs:377-455; returns when the game ends or the player quits."""
    stdin = stdin or sys.stdin
    stdout = stdout or sys.stdout

    def say(text="", end="\n"):
        stdout.write(text + end)
        stdout.flush()

    while True:
        if bound is None:
            _, premonitions = engine.lookahead(world)
            if not len(premonitions):
                return say("THE END")
            say(depict_world(world, colour))
            for index, (premonition, plus) in enumerate(zip(premonitions, _endangering(engine, premonitions))):
                say("%2d. %s" % (index, depict_commit(premonition, plus, colour)))
        else:
            forecasts, depth, thinking_time = co.forecast(engine, world, bound)
            say(depict_world(world, colour))
            report = str(depth)
            if isinstance(bound, co.Depth) and bound.extension is not None:
                report = "at least %d and up to %d" % (bound.depth, bound.depth + bound.extension)
            say("(scoring alternatives %s levels deep took %d ms)" % (report, thinking_time))
            premonitions = [f["commit"] for f in forecasts]
            if not len(premonitions):
                return say("THE END")
            for index, (sight, plus) in enumerate(zip(forecasts, _endangering(engine, premonitions))):
                score = "%.1f" % float(sight["score"])
                if colour:
                    score = "\x1b[1;35m" + score + _RESET
                say("%2d: %s — score %s ‣ representative variation: %s" % (index, depict_commit(sight["commit"], plus, colour), score,
                                                                           variation_format(sight)))
        while True:
            say("\nSelect a move>> ", end="")
            line = stdin.readline()
            if not line or line.strip() == "quit":
                return say("THE END")
            try:
                choice = int(line.strip())
                if choice < 0:
                    raise ValueError("negative")
            except ValueError as error:
                say("Error parsing choice: %r. Try again." % (error,))
                continue
            if choice < len(premonitions):
                world = premonitions[choice]["tree"].copy()
                break
            say("%d isn't among the choices. Try again." % choice)


def bound_from_args(depth, quiet, depth_sequence, seconds):
    """LookaheadBound::from_args (main.rs:119-150) -> a bound, None, or ValueError"""
    given = [b for b in (depth, depth_sequence, seconds) if b is not None]
    if len(given) > 1:
        raise ValueError("more than one of `--depth`, `--depth-sequence`, or `--seconds` was passed")
    if depth is not None:
        return co.Depth(depth, quiet)
    if depth_sequence is not None:
        return co.DepthSequence.from_depiction(depth_sequence)
    if seconds is not None:
        return co.Seconds(seconds)
    return None


def game(engine, from_runes=None, bound=None, stdin=None, stdout=None, colour=True):
    out = stdout or sys.stdout
    out.write("Welcome to Leafline on B200!\n")
    world = runes.reconstruct(from_runes) if from_runes else runes.reconstruct("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq -")
    play(engine, world, bound, stdin, stdout, colour)
