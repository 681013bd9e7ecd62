#!/usr/bin/env python3
"""Regenerates tests/golden/*.json|txt from the reference's OWN test files (run in the build container only:
/root/reference does not exist on the GPU box).  Only known answers are extracted -- PTN move lists,
expected counts/results/strings -- never code.

  python tests/golden/make_golden.py [/root/reference]
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def read(rel):
    with open(os.path.join(REF, rel)) as f:
        return f.read()


def str_list(block):
    return re.findall(r'"([^"]*)"', block)


def num(text):
    return int(text.replace("_", ""))


def perft_golden():
    src = read("tak/tests/perft.rs")
    cases = []
    # positional perfts: fn name, board size, PTN list, asserts
    for m in re.finditer(r"fn (\w+)\(\)[^{]*\{(.*?)\n\}", src, re.S):
        name, body = m.group(1), m.group(2)
        gm = re.search(r"Game::<(\d)>::from_ptn_moves\(&\[(.*?)\]\)", body, re.S)
        if gm:
            n = int(gm.group(1))
            moves = str_list(gm.group(2))
            exp = [(int(d), num(v)) for d, v in re.findall(r"^\s*assert_eq!\(perf_count\(&game, (\d+)\), ([\d_]+)\);", body, re.M)]
            commented = [(int(d), num(v)) for d, v in re.findall(r"//\s*assert_eq!\(perf_count\(&game, (\d+)\), ([\d_]+)\);", body)]
            cases.append({"name": name, "n": n, "moves": moves, "expect": exp, "commented": commented})
        else:
            exp = re.findall(r"^\s*assert_eq!\(perf_count\(&Game::<(\d)>::default\(\), (\d+)\), ([\d_]+)\);", body, re.M)
            com = re.findall(r"//\s*assert_eq!\(perf_count\(&Game::<(\d)>::default\(\), (\d+)\), ([\d_]+)\);", body)
            if exp:
                cases.append({"name": name, "n": int(exp[0][0]), "moves": [],
                              "expect": [(int(d), num(v)) for _, d, v in exp],
                              "commented": [(int(d), num(v)) for _, d, v in com]})
    return cases


def wins_golden():
    src = read("tak/tests/wins.rs")
    cases = []
    for m in re.finditer(r"fn (\w+)\(\)[^{]*\{(.*?)\n\}", src, re.S):
        name, body = m.group(1), m.group(2)
        gm = re.search(r"Game::<(\d)>::from_ptn_moves\(&\[(.*?)\]\)", body, re.S)
        n, moves = int(gm.group(1)), str_list(gm.group(2))
        checks = []
        # sequence of (optional half_komi assignment, expected result)
        pos = 0
        half_komi = None
        for ev in re.finditer(r"game\.half_komi = (\d+);|assert_eq!\(game\.result\(\), GameResult::(\w+) \{(.*?)\}\)", body, re.S):
            if ev.group(1) is not None:
                half_komi = int(ev.group(1))
            else:
                kind, fields = ev.group(2), ev.group(3)
                if kind == "Winner":
                    color = re.search(r"Color::(\w+)", fields).group(1)
                    road = "true" in re.search(r"road:\s*(\w+)", fields).group(1)
                    code = (1 if color == "White" else 2) | (0x10 if road else 0)
                else:
                    rev = "true" in re.search(r"reversible_plies:\s*(\w+)", fields).group(1)
                    code = 3 | (0x10 if rev else 0)
                checks.append({"half_komi": half_komi, "result": code})
        cases.append({"name": name, "n": n, "moves": moves, "checks": checks})
    return cases


def tps_golden():
    src = read("tak/tests/tps.rs")
    m = re.search(r"fn complicated_board\(\) \{(.*?)\n\}", src, re.S)
    body = m.group(1)
    gm = re.search(r"Game::<(\d)>::from_ptn_moves\(&\[(.*?)\]\)", body, re.S)
    n, moves = int(gm.group(1)), str_list(gm.group(2))
    tm = re.search(r'tps\.to_string\(\),\s*"(.*?)"\s*\)', body, re.S)
    tps = re.sub(r"\\\n\s*", "", tm.group(1))
    seeds = [int(s) for s in re.findall(r"tps_consistency\((\d+)\)\n", src)]
    return {"n": n, "moves": moves, "tps": tps, "consistency_seeds": seeds, "consistency_n": 5}


def symm_seeds():
    src = read("tak/tests/symm.rs")
    return [int(s) for s in re.findall(r"symmetrical_boards\((\d+)\)\n", src)]


def repr_golden():
    src = read("alpha-tak/src/repr/tests.rs")
    m = re.search(r"fn complicated_board\(\) \{(.*?)\n\}", src, re.S)
    body = m.group(1)
    gm = re.search(r"Game::<(\d)>::from_ptn_moves\(&\[(.*?)\]\)", body, re.S)
    n, moves = int(gm.group(1)), str_list(gm.group(2))
    sm = re.search(r"Tensor::of_slice\(&\[(.*?)\]\)\.view\(\[12, 5, 5\]\)", body, re.S)
    cells = re.findall(r"\b([xo])\b,", re.sub(r"//.*", "", sm.group(1)))
    assert len(cells) == 12 * 25, len(cells)
    planes = [1 if c == "x" else 0 for c in cells]
    return {"n": n, "moves": moves, "to_move_arg": 0, "planes_12x5x5": planes}


def mcts_golden():
    src = read("alpha-tak/src/search/tests.rs")
    out = []
    for m in re.finditer(r"fn (\w+)\(\) \{(.*?)\n\}", src, re.S):
        name, body = m.group(1), m.group(2)
        gm = re.search(r"Game::<(\d)>::from_ptn_moves\(&\[(.*?)\]\)", body, re.S)
        if not gm:
            continue
        out.append({"name": name, "n": int(gm.group(1)), "moves": str_list(gm.group(2)),
                    "rollouts": int(re.search(r"0\.\.(\d+)", body).group(1))})
    return out


def move_table_5():
    src = read("alpha-tak/src/search/move_map.rs")
    m = re.search(r"const POSSIBLE_MOVES_IN_5S: \[&str; 1575\] = \[(.*?)\];", src, re.S)
    lst = str_list(m.group(1))
    assert len(lst) == 1575
    return lst


def main():
    golden = {
        "source": "ViliamVadocz/tak tests (tak/tests/{perft,wins,tps,symm}.rs, alpha-tak/src/{repr,search}/tests.rs)",
        "perft": perft_golden(),
        "wins": wins_golden(),
        "tps": tps_golden(),
        "symm_seeds": symm_seeds(),
        "board_repr": repr_golden(),
        "mcts": mcts_golden(),
    }
    with open(os.path.join(OUT, "reference_known_answers.json"), "w") as f:
        json.dump(golden, f, indent=1)
    with open(os.path.join(OUT, "move_index_5.txt"), "w") as f:
        f.write("\n".join(move_table_5()) + "\n")
    print("perft cases:", [(c["name"], c["expect"]) for c in golden["perft"]])
    print("wins:", [(c["name"], c["checks"]) for c in golden["wins"]])
    print("tps:", golden["tps"]["tps"], golden["tps"]["consistency_seeds"])
    print("mcts:", golden["mcts"])


if __name__ == "__main__":
    main()
