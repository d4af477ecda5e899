"""The reference's command-line tool (src/main.rs) rebuilt over the C ABI (cli/kzg_cli.cpp): same subcommands,
artifact files, JSON bytes and error messages.  The CPU tests cover argument handling and the errors raised before
any bulk work; the GPU test runs the whole ceremony -> commit -> evaluate -> verify flow and compares every artifact
byte for byte with what serde_json would write for the oracle's values."""
import json
import os
import subprocess

import pytest

import kzg_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cli", "kzg-cli")
CPP_TEST_SRC = os.path.join(ROOT, "tests", "cpp", "test_api.cpp")
CPP_TEST_BIN = os.path.join(ROOT, "tests", "cpp", "test_api.bin")


@pytest.fixture(scope="module")
def cli(native):
    from kzg_poly_commit_exploration_b200 import build

    build.build_cli()
    assert os.path.exists(CLI)
    return CLI


def run(cli, args, cwd, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([cli, *args], cwd=cwd, capture_output=True, text=True, env=e, timeout=600)


def compact(obj) -> str:
    return json.dumps(obj, separators=(",", ":"))


def test_cli_argument_handling_and_early_errors(cli, tmp_path):
    r = run(cli, [], tmp_path)
    assert r.returncode == 0 and "No command has been input" in r.stdout            # src/main.rs:82-84
    r = run(cli, ["--version"], tmp_path)
    assert r.stdout.strip() == "kzg_poly_commit_exploration 0.1.0"
    r = run(cli, ["commit"] + [str(i) for i in range(1, 12)], tmp_path)               # degree 10 > MAX_DEGREE
    assert r.returncode == 101
    assert "Command execution failed with error: Unhandled error: Only polynomials up to degree 9 are supported" in r.stderr
    r = run(cli, ["commit", "1", "2", "0", "0"], tmp_path)                            # trimmed to degree 1; no setup yet
    assert r.returncode == 101 and 'P(x) = "2x + 1"' in r.stdout
    assert "Trusted setup artifacts have not been found, generate them beforehand." in r.stderr
    r = run(cli, ["commit", "--", "-1", "5"], tmp_path)                               # negative i128: r - 1
    assert f'"5x + {O.R - 1}"' in r.stdout
    r = run(cli, ["evaluate", "3"], tmp_path)
    assert r.returncode == 101 and "Trusted setup artifacts have not been found" in r.stderr
    r = run(cli, ["commit", "abc"], tmp_path)
    assert r.returncode == 2 and "invalid value 'abc'" in r.stderr
    r = run(cli, ["commit", str(1 << 127)], tmp_path)                                 # out of i128 range
    assert r.returncode == 2
    r = run(cli, ["nope"], tmp_path)
    assert r.returncode == 2 and "unrecognized subcommand" in r.stderr


def test_cpp_api_cpu_part(native):
    """include/kzg_b200.hpp: Scalar / G1Point mirror against published vectors (runs without a GPU)."""
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), CPP_TEST_SRC, "-o", CPP_TEST_BIN,
                    "-L", os.path.join(ROOT, "kzg_poly_commit_exploration_b200"), "-l:libkzg_b200.so",
                    "-Wl,-rpath," + os.path.join(ROOT, "kzg_poly_commit_exploration_b200")], check=True)
    r = subprocess.run([CPP_TEST_BIN], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "FAIL" not in r.stdout


@pytest.mark.gpu
def test_cpp_api_full(native):
    test_cpp_api_cpu_part(native)
    r = subprocess.run([CPP_TEST_BIN], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "C++ API test ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cli_full_flow_matches_serde_json_of_oracle_values(cli, tmp_path):
    env = {"KZG_CLI_SECRET_HEX": O.BENCH_SECRET_BE.hex()}
    s = O.fr_from_be_bytes(O.BENCH_SECRET_BE)
    r = run(cli, ["trusted-setup"], tmp_path, env)
    assert r.returncode == 0, r.stderr
    setup_json = (tmp_path / "artifacts" / "setup.json").read_text()
    expected = [{"g1": list(O.g1_compress(O.g1_from_scalar(pow(s, i, O.R)))),
                 "g2": list(O.g2_compress(O.g2_from_scalar(pow(s, i, O.R))))} for i in range(10)]
    assert setup_json == compact(expected)
    # commit: the bench polynomial c_i = 5^i + 10 of degree 9, plus two trailing zeros that must be trimmed
    co = O.bench_polynomial(9)
    r = run(cli, ["commit"] + [str(c) for c in co] + ["0", "0"], tmp_path)
    assert r.returncode == 0, r.stderr
    commitment = O.g1_compress(O.commit_known_secret(co, s))
    assert (tmp_path / "artifacts" / "commitment.json").read_text() == compact(
        {"polynomial": {"coefficients": [list(O.fr_to_le_bytes(c)) for c in co]}, "commitment": list(commitment)})
    # evaluate at a negative point
    x = -12345678901234567890
    z = O.fr_from_i128(x)
    y = O.horner(co, z)
    r = run(cli, ["evaluate", "--", str(x)], tmp_path)
    assert r.returncode == 0, r.stderr
    assert f'with "P({x}) = {y}"' in r.stdout
    proof = O.g1_compress(O.proof_known_secret(co, s, z))
    assert (tmp_path / "artifacts" / "evaluation.json").read_text() == compact(
        {"evaluation": {"point": list(O.fr_to_le_bytes(z)), "result": list(O.fr_to_le_bytes(y))}, "proof": list(proof)})
    r = run(cli, ["verify-evaluation"], tmp_path)
    assert r.returncode == 0 and "Successfully verified evaluation" in r.stdout, r.stderr
    # a tampered evaluation is rejected (src/main.rs:290-295)
    ev = json.loads((tmp_path / "artifacts" / "evaluation.json").read_text())
    ev["evaluation"]["result"][0] ^= 1
    (tmp_path / "artifacts" / "evaluation.json").write_text(compact(ev))
    r = run(cli, ["verify-evaluation"], tmp_path)
    assert r.returncode == 101 and "The proof associated to the evaluation is incorrect." in r.stderr
    # a malformed point in the setup is a deserialisation error
    st = json.loads(setup_json)
    st[3]["g1"][0] &= 0x7F
    (tmp_path / "artifacts" / "setup.json").write_text(compact(st))
    r = run(cli, ["commit", "1", "2"], tmp_path)
    assert r.returncode == 101 and "uncompressing" in r.stderr
