"""The front ends either side of the leaf layer -- console game, UCI dæmon, `--correspond` wire format -- against a
stand-in engine made of the oracle.  No GPU: the front ends compute nothing themselves (the same dialogues run against
the real engine in tests/test_gpu_parity.py)."""
import io

import numpy as np
import pytest

import oracle as O
from leafline_b200 import console, correspondence as co, uci
from leafline_b200._abi import FORECAST_DTYPE


class OracleEngine:
    """The three calls the console makes, answered by the oracle."""

    def lookahead(self, world):
        commits = O.lookahead(np.ascontiguousarray(world))
        return np.array([0, len(commits)], dtype=np.uint32), commits

    def in_critical_endangerment(self, worlds, teams):
        teams = np.broadcast_to(np.asarray(teams), (len(worlds),))  # a scalar team applies to every world, like Engine's
        return np.array([bool(O.in_critical_endangerment(np.ascontiguousarray(w), int(t))) for w, t in zip(worlds, teams)])

    def deep_forecast(self, world, depth, nihilistically=False, quiet=None):
        roots, scores = O.kickoff(np.ascontiguousarray(world), depth, nihilistically, quiet=quiet)
        order = sorted(range(len(roots)), key=lambda i: -scores[i])
        out = np.zeros(len(roots), dtype=FORECAST_DTYPE)
        for k, i in enumerate(order):
            out[k]["commit"] = roots[i]
            out[k]["score"] = scores[i]
            out[k]["variation_len"] = 1
            for field in ("team", "job", "whence", "whither"):
                out[k]["variation"][0][field] = roots[i][field]
        return out

    forecast = deep_forecast

    def fixed_depth_sequence_kickoff(self, world, depths, nihilistically=False):
        return self.deep_forecast(world, depths[-1], nihilistically)

    def iterative_deepening_kickoff(self, world, seconds, nihilistically=False):
        return self.deep_forecast(world, 2, nihilistically), 2


def test_world_depiction():
    text = console.depict_world(O.new_world(), colour=False)
    lines = text.splitlines()
    assert lines[0] == "    a b c d e f g h"
    assert lines[1] == " 8 ♜ ♞ ♝ ♛ ♚ ♝ ♞ ♜ " and lines[2] == " 7 " + "♟ " * 8
    assert lines[3] == " 6 " + "  " * 8 and lines[8] == " 1 ♜ ♞ ♝ ♛ ♚ ♝ ♞ ♜ "  # solid runes for both teams (identity.rs:179-184)
    painted = console.depict_world(O.new_world(), colour=True)
    assert "\x1b[48;5;58m" in painted and "\x1b[48;5;17m" in painted and "\x1b[1;31m" in painted and "\x1b[36m" in painted


def test_movement_runes():
    eng = OracleEngine()
    # a stun with ascension, giving check: life.rs's own rune format (prefix x locale + =)
    w = O.reconstruct("1n2k3/P7/8/8/8/8/8/4K3 w - -")
    _, commits = eng.lookahead(w)
    plus = eng.in_critical_endangerment([c["tree"] for c in commits], [int(c["tree"]["initiative"]) for c in commits])
    runes = [console.pagan_movement_rune(c, p) for c, p in zip(commits, plus)]
    assert runes[:8] == ["a8=N", "a8=B", "a8=R", "a8=Q", "xb8=N", "xb8=B", "xb8+=R", "xb8+=Q"]  # the pony on b8 shields the figurehead from a8
    text = [console.depict_commit(c, p, colour=False) for c, p in zip(commits, plus)]
    assert "xb8+=Q (♙ from a7 to b8, stunning ♞, transitioning into ♕)" in text
    assert "a8=N (♙ from a7 to a8, transforming into ♘)" in text and "xb8+=R (♙ from a7 to b8, stunning ♞, being brevetted to ♖)" in text
    assert "Kd1 (♔ from e1 to d1)" in text
    # secret service and Blue's ".." prefix
    w = O.reconstruct("r3k2r/8/8/8/8/8/8/4K3 b kq -")
    _, commits = eng.lookahead(w)
    text = [console.depict_commit(c, False, colour=False) for c in commits]
    assert "..O–O (♚ from e8 to g8)" in text and "..O–O–O (♚ from e8 to c8)" in text and "..Ra7 (♜ from a8 to a7)" in text


def test_the_prompt_loop_without_a_bound():
    eng = OracleEngine()
    out = io.StringIO()
    console.play(eng, O.new_world(), None, io.StringIO("12\nfoo\n99\n0\nquit\n"), out, colour=False)
    text = out.getvalue()
    assert text.count("    a b c d e f g h") == 3  # the opening, after choice 12, after choice 0
    assert " 0. a3 (♙ from a2 to a3)" in text and "19. Nh3 (♘ from g1 to h3)" in text
    assert " 0. ..a6 (♟ from a7 to a6)" in text  # Blue's reply list follows Orange's choice
    assert "Error parsing choice: " in text and "99 isn't among the choices. Try again." in text
    assert text.count("Select a move>> ") == 5 and text.rstrip().endswith("THE END")


def test_the_prompt_loop_with_forecasts_and_the_end():
    eng = OracleEngine()
    out = io.StringIO()
    # mind.rs:662-675: the under-ascension to a pony wins the princess; depth 3 from there
    console.play(eng, O.reconstruct("8/q1P1k/8/8/8/8/6PP/7K w - -"), co.Depth(3, None), io.StringIO("0\nquit\n"), out, colour=False)
    text = out.getvalue()
    assert "(scoring alternatives 3 levels deep took " in text
    assert " 0: c8+=N (♙ from c7 to c8, transforming into ♘) — score 4.7 ‣ representative variation: c8" in text
    assert "..Kd" in text or "..Ke" in text or "..Kf" in text  # Blue must answer the check
    # a world with nothing to play ends the game at once
    out = io.StringIO()
    console.play(eng, O.reconstruct("7k/5Q2/6K1/8/8/8/8/8 b - -"), None, io.StringIO(""), out, colour=False)
    assert out.getvalue() == "THE END\n"
    out = io.StringIO()
    console.play(eng, O.reconstruct("8/q1P1k/8/8/8/8/6PP/7K w - -"), co.Depth(1, 2), io.StringIO("quit\n"), out, colour=False)
    assert "(scoring alternatives at least 1 and up to 3 levels deep took " in out.getvalue()


def test_bounds_from_arguments():
    assert console.bound_from_args(None, None, None, None) is None
    assert console.bound_from_args(4, 2, None, None).extension == 2
    assert console.bound_from_args(None, None, "1,2,3", None).depths == [1, 2, 3]
    assert console.bound_from_args(None, None, None, 5).seconds == 5
    with pytest.raises(ValueError):
        console.bound_from_args(4, None, None, 5)


def test_uci_dialogue():
    # uci.rs:13-104
    script = "uci\nisready\nucinewgame\nposition startpos moves e2e4\ngo depth 2\nposition startpos moves e2e4 x g1f3\ngo wtime 4000 btime 4000 movestogo 2\nquit\n"
    out = io.StringIO()
    uci.dæmon(engine=OracleEngine(), stdin=io.StringIO(script), stdout=out)
    lines = out.getvalue().splitlines()
    assert lines[0].startswith("id name Leafline") and lines[1].startswith("id author") and lines[2] == "uciok" and lines[3] == "readyok"
    best = [l for l in lines if l.startswith("bestmove ")]
    assert len(best) == 2 and all(len(b.split()[1]) == 4 for b in best)
    assert any(l.startswith("info depth 2 score cp ") for l in lines)
    # the reply to 1.e4 at depth 2 is the move exact negamax ranks first (stable order)
    after_e4 = [c for c in O.lookahead(O.new_world()) if O.to_algebraic(c["whence"]) + O.to_algebraic(c["whither"]) == "e2e4"][0]["tree"]
    roots, scores = O.kickoff(np.ascontiguousarray(after_e4), 2, False)
    top = roots[int(np.argmax(scores))]
    assert best[0] == "bestmove " + O.to_algebraic(top["whence"]) + O.to_algebraic(top["whither"])
    with pytest.raises(ValueError):
        uci.dæmon(engine=OracleEngine(), stdin=io.StringIO("frobnicate\n"), stdout=io.StringIO())


def test_correspondence_cards():
    import json

    eng = OracleEngine()
    # main.rs:461-469: the reference's own end-to-end test
    assert co.correspondence(eng, "R6k/6pp/8/8/8/8/8/8 b - -", co.Depth(2, None)) == '{"the_triumphant":"Orange"}'
    assert co.correspondence(eng, "7k/5Q2/6K1/8/8/8/8/8 b - -", co.Depth(1, None)) == '{"the_triumphant":null}'
    card = co.correspondence(eng, "8/q1P1k/8/8/8/8/6PP/7K w - -", co.DepthSequence.from_depiction("1,2,3"))
    d = json.loads(card)
    assert list(d) == ["world", "patch", "hospitalization", "thinking_time", "depth", "counterreplies", "rosetta_stone"]
    assert d["world"] == "2N5/q3k3/8/8/8/8/6PP/7K b - -" and d["depth"] == 3 and d["hospitalization"] is None and d["rosetta_stone"] == "c8"
    assert d["patch"] == {"star": {"team": "Orange", "job_description": "Servant"}, "whence": {"rank": 6, "file": 2}, "whither": {"rank": 7, "file": 2}}
    assert len(d["counterreplies"]) == len(O.lookahead(O.reconstruct(d["world"]))) and ", " not in card and "\": " not in card
