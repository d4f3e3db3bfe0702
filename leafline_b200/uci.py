"""UCI front end over the GPU leaf layer -- the protocol subset the reference's dæmon speaks (src/uci.rs:13-104), so that
a GUI configured for the reference binary can be pointed at `python -m leafline_b200 --uci` unchanged.

What is kept of the reference's behaviour, because GUIs depend on it: the identification lines, `readyok`, one `bestmove`
per `go`, an `info depth .. score cp ..` line before the move of a clock-bound search, the engine's own reply applied to its
position, and `position ... moves ...` read incrementally -- the position the dæmon already holds is trusted and only the
LAST move of the line is matched against `lookahead()` (whence + whither; an unmatched move changes nothing).  What differs is
the engine behind it: `go depth N` is `Engine.deep_forecast` (full-width on the device, the host alpha-beta driver beyond the
HBM limit), a clock-bound `go` is `Engine.iterative_deepening_kickoff`.

Structure: a `Daemon` object holding the position, one method per command, a dispatch table.
"""
import sys

from . import Engine, algebraic, new_world, to_algebraic

VERSION = "0.0.18-b200"  # the reference crate is 0.0.18 (Cargo.toml)
AUTHOR = "Zack M. Davis and friends"


class Quit(Exception):
    """raised by the `quit` command to leave the read loop"""


class Daemon:
    def __init__(self, engine, stdout):
        self.engine = engine
        self.stdout = stdout
        self.world = new_world()
        self.commands = {
            "uci": self.identify,
            "isready": lambda _: self.say("readyok"),
            "ucinewgame": self.new_game,
            "quit": self.quit,
            "position": self.position,
            "go": self.go,
        }

    def say(self, line):
        self.stdout.write(line + "\n")
        self.stdout.flush()

    # ---- commands
    def identify(self, _):
        self.say(f"id name Leafline v. {VERSION}")
        self.say(f"id author {AUTHOR}")
        self.say("uciok")

    def new_game(self, _):
        self.world = new_world()

    def quit(self, _):
        raise Quit

    def position(self, arguments):
        if not arguments:
            raise ValueError("expected a discernable last move")
        self.play(arguments[-1])

    def go(self, arguments):
        if len(arguments) % 2:
            raise ValueError("expected option value")
        options = {key: int(value) for key, value in zip(arguments[0::2], arguments[1::2])}  # ValueError: "expected an integer literal"
        if "depth" in options:
            forecasts = self.engine.deep_forecast(self.world, options["depth"])
        else:
            forecasts, depth = self.engine.iterative_deepening_kickoff(self.world, self.seconds_for_this_move(options))
            self.say(f"info depth {depth} score cp {int(float(forecasts[0]['score']) * 100.0)}")
        reply = forecasts[0]["commit"]
        self.say(f"bestmove {to_algebraic(reply['whence'])}{to_algebraic(reply['whither'])}")
        self.world = reply["tree"].copy()

    # ---- helpers
    def play(self, move):
        """apply `move` (e.g. "e2e4") if it names a legal commit of the held position"""
        whence, whither = algebraic(move[0:2]), algebraic(move[2:4])
        _, commits = self.engine.lookahead(self.world)
        for commit in commits:
            if commit["whence"] == whence and commit["whither"] == whither:
                self.world = commit["tree"].copy()  # (ascensions share a patch: the last one listed wins, as in the reference)

    def seconds_for_this_move(self, options):
        """the reference's time management: an even share of the mover's clock plus the increment, in whole seconds"""
        orange = int(self.world["initiative"]) == 0
        clock, increment = options["wtime" if orange else "btime"], options.get("winc" if orange else "binc", 0)
        return (clock // options["movestogo"] + increment) // 1000

    def handle(self, line):
        words = line.split()
        if not words:
            raise ValueError("expected a command")
        command = self.commands.get(words[0])
        if command is None:
            raise ValueError(f"got unrecognized UCI command {words[0]!r}")  # the reference panics here (uci.rs:100)
        command(words[1:])


def dæmon(engine=None, stdin=None, stdout=None):
    """read UCI commands from `stdin` until `quit` (or end of input)"""
    stdin, stdout = stdin or sys.stdin, stdout or sys.stdout
    own = engine is None
    engine = engine or Engine(0)
    daemon = Daemon(engine, stdout)
    try:
        for line in stdin:
            try:
                daemon.handle(line)
            except Quit:
                break
    finally:
        if own:
            engine.close()
