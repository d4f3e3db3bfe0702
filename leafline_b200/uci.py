"""Unlikely Command Integration dæmon (src/uci.rs:13-104) over the GPU leaf layer.

Same protocol subset and the same shortcuts as the reference: `uci`, `isready`, `ucinewgame`, `quit`,
`position ... <last move>` (only the last move of the line is looked at and matched against lookahead()),
`go depth N` / `go wtime .. btime .. movestogo .. [winc ..] [binc ..]`.
"""
import sys

from . import Engine, algebraic, new_world, to_algebraic

VERSION = "0.0.18-b200"  # the reference crate is 0.0.18 (Cargo.toml)


def dæmon(engine=None, stdin=None, stdout=None):
    stdin, stdout = stdin or sys.stdin, stdout or sys.stdout
    own = engine is None
    engine = engine or Engine(0)
    world = new_world()

    def say(line):
        stdout.write(line + "\n")
        stdout.flush()

    try:
        for line in stdin:
            if line == "uci\n":
                say(f"id name Leafline v. {VERSION}")
                say("id author Zack M. Davis and friends")
                say("uciok")
                continue
            if line == "isready\n":
                say("readyok")
                continue
            if line == "ucinewgame\n":
                world = new_world()
                continue
            if line == "quit\n":
                break
            tokens = line.split()
            if not tokens:
                raise ValueError("expected a command")
            command, tokens = tokens[0], tokens[1:]
            if command == "position":
                # uci.rs:47-62: trust the internal WorldState and look only at the last-move diff
                if not tokens:
                    raise ValueError("expected a discernable last move")
                diff = tokens[-1]
                whence, whither = algebraic(diff[0:2]), algebraic(diff[2:4])
                _, premonitions = engine.lookahead(world)
                for premonition in premonitions:
                    if premonition["whence"] == whence and premonition["whither"] == whither:
                        world = premonition["tree"].copy()
            elif command == "go":
                options = {}
                it = iter(tokens)
                for key in it:
                    options[key] = int(next(it))  # "expected option value" / "expected an integer literal"
                if "depth" in options:
                    forecasts = engine.deep_forecast(world, options["depth"])
                else:
                    time_key, increment_key = ("wtime", "winc") if int(world["initiative"]) == 0 else ("btime", "binc")
                    deadline = (options[time_key] // options["movestogo"] + options.get(increment_key, 0)) // 1000
                    forecasts, depth = engine.iterative_deepening_kickoff(world, deadline)
                    say(f"info depth {depth} score cp {int(float(forecasts[0]['score']) * 100.0)}")
                movement = forecasts[0]["commit"]
                say(f"bestmove {to_algebraic(movement['whence'])}{to_algebraic(movement['whither'])}")
                world = movement["tree"].copy()
            else:
                raise ValueError(f"got unrecognized UCI command {command!r}")  # moral_panic!, uci.rs:100
    finally:
        if own:
            engine.close()
