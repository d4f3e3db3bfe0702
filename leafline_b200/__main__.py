"""`python -m leafline_b200` -- the reference binary's entry points (src/main.rs:253-456):

    python -m leafline_b200 --correspond --depth 4 --from "<preservation runes>"     # what the web client shells out to
    python -m leafline_b200 --correspond --seconds 5 --from "<runes>"
    python -m leafline_b200 --uci
    python -m leafline_b200 [--depth 4 | --seconds 5 | --depth-sequence 2,4] [--from "<runes>"]   # the console game

`--déjà-vu-bound` and `--debug` are accepted and ignored (the search's memory bank sizes itself, there is no log file).
"""
import argparse
import sys

from . import Engine, console, correspondence as co, uci


def main(argv=None):
    ap = argparse.ArgumentParser(prog="leafline_b200", description="Leafline: an oppositional strategy game engine (B200 leaf layer)")
    ap.add_argument("--depth", type=int, help="rank moves using AI minimax lookahead this deep")
    ap.add_argument("--quiet", type=int, help="search with quietness extension this deep")
    ap.add_argument("--depth-sequence", help="rank moves using AI minimax lookahead to these depths")
    ap.add_argument("--seconds", type=int, help="rank moves using AI minimax for about this many seconds")
    ap.add_argument("--correspond", action="store_true", help="just output the serialization of the AI's top response and legal replies thereto")
    ap.add_argument("--uci", "--deamon", "--dæmon", dest="uci", action="store_true", help="run Unlikely Command Integration dæmon for external driver play")
    ap.add_argument("--from", dest="from_runes", help="start a game from the given book of preservation runes")
    ap.add_argument("--déjà-vu-bound", "--deja-vu-bound", dest="deja_vu_bound", type=float, default=2.0, help="ignored: there is no déjà vu table on the device")
    ap.add_argument("--debug", action="store_true", help="ignored")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)
    if args.correspond:
        # LookaheadBound::from_args, main.rs:119-150
        given = [b for b in (args.depth, args.depth_sequence, args.seconds) if b is not None]
        if len(given) > 1:
            sys.exit("more than one of `--depth`, `--depth-sequence`, or `--seconds` was passed")
        if not given:
            sys.exit("`--correspond` passed without exactly one of `--depth`, `--depth-sequence`, or `--seconds`")
        if args.from_runes is None:
            sys.exit("`--correspond` requires `--from`")
        if args.depth is not None:
            bound = co.Depth(args.depth, args.quiet)
        elif args.depth_sequence is not None:
            bound = co.DepthSequence.from_depiction(args.depth_sequence)
        else:
            bound = co.Seconds(args.seconds)
        with Engine(args.device) as engine:
            print(co.correspondence(engine, args.from_runes, bound))
        return 0
    if args.uci:
        uci.dæmon()
        return 0
    try:
        bound = console.bound_from_args(args.depth, args.quiet, args.depth_sequence, args.seconds)
    except ValueError as error:
        sys.exit(str(error))
    with Engine(args.device) as engine:
        console.game(engine, args.from_runes, bound, colour=sys.stdout.isatty())
    return 0


if __name__ == "__main__":
    sys.exit(main())
