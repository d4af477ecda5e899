// kzg-cli — the reference's command-line tool (src/main.rs) on top of the B200 engine.
//
// Same four subcommands, same artifact files and byte-identical JSON (serde_json's compact output of
// serialize_bytes: arrays of integers; field order g1,g2 / polynomial,commitment / evaluation,proof),
// same MAX_DEGREE = 9 check and error messages:
//   trusted-setup        -> ./artifacts/setup.json       Vec<SetupArtifact>            src/main.rs:112-140
//   commit <c0> <c1> ..  -> ./artifacts/commitment.json  CommitmentArtifact            src/main.rs:141-188
//   evaluate <x>         -> ./artifacts/evaluation.json  EvaluationArtifact            src/main.rs:189-243
//   verify-evaluation                                                                  src/main.rs:244-305
// Every bulk operation is one C-ABI call into libkzg_b200.so (there is no CPU fallback: without a CUDA device
// the tool fails like any other command error).  A failed command ends like the reference's panic: the message
// "Command execution failed with error: Unhandled error: ..." on stderr and exit status 101.
// Test hook (not in the reference): KZG_CLI_SECRET_HEX = 64 hex digits replaces the random secret.
//
// Build: g++ -std=c++17 -O2 -Iinclude cli/kzg_cli.cpp -o cli/kzg-cli kzg_poly_commit_exploration_b200/libkzg_b200.so \
//            -Wl,-rpath,'$ORIGIN/../kzg_poly_commit_exploration_b200'
#include <sys/stat.h>
#include <unistd.h>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <sstream>
#include "kzg_b200.hpp"

using namespace kzg;
using curves::G1Point;
using curves::G2Point;
using polynomial::Evaluation;
using polynomial::Polynomial;
using scalar::Scalar;
using trusted_setup::SetupArtifact;
using trusted_setup::TrustedSetup;

static const char* ARTIFACTS_FOLDER_PATH = "./artifacts";
static const char* SETUP_ARTIFACTS_PATH = "./artifacts/setup.json";
static const char* COMMITMENT_ARTIFACTS_PATH = "./artifacts/commitment.json";
static const char* EVALUATION_ARTIFACTS_PATH = "./artifacts/evaluation.json";
static const uint32_t MAX_DEGREE = 9;

struct CliError : std::runtime_error { using std::runtime_error::runtime_error; };

// ---- logging (simple_logger's line shape) -------------------------------------------------------------
static int g_level = 3;  // 1 error, 2 warn, 3 info, 4 debug, 5 trace
static void log_line(int level, const char* name, const std::string& msg) {
  if (level > g_level) return;
  char ts[32]; time_t t = time(nullptr); struct tm tmv; gmtime_r(&t, &tmv);
  strftime(ts, sizeof ts, "%Y-%m-%dT%H:%M:%SZ", &tmv);
  printf("%s %-5s [kzg_poly_commit_exploration] %s\n", ts, name, msg.c_str());
  fflush(stdout);
}
static void info(const std::string& m) { log_line(3, "INFO", m); }
static void warn(const std::string& m) { log_line(2, "WARN", m); }

// ---- a JSON reader for the three artifact shapes (objects, arrays, unsigned integers, null) ----------------
struct Json {
  enum Kind { Null, Num, Arr, Obj } kind = Null;
  uint64_t num = 0;
  std::vector<Json> items;                                // Arr
  std::vector<std::pair<std::string, Json>> fields;       // Obj
  const Json& at(const std::string& k) const {
    for (auto& f : fields) if (f.first == k) return f.second;
    throw CliError("missing field `" + k + "`");
  }
};
struct JsonParser {
  const std::string& s; size_t i = 0;
  explicit JsonParser(const std::string& src) : s(src) {}
  void ws() { while (i < s.size() && isspace((unsigned char)s[i])) i++; }
  [[noreturn]] void fail(const char* what) { throw CliError(std::string(what) + " at column " + std::to_string(i + 1)); }
  Json value() {
    ws();
    if (i >= s.size()) fail("EOF while parsing a value");
    Json j;
    char c = s[i];
    if (c == '[') {
      j.kind = Json::Arr; i++; ws();
      if (i < s.size() && s[i] == ']') { i++; return j; }
      for (;;) { j.items.push_back(value()); ws(); if (i < s.size() && s[i] == ',') { i++; continue; } if (i < s.size() && s[i] == ']') { i++; break; } fail("expected `,` or `]`"); }
    } else if (c == '{') {
      j.kind = Json::Obj; i++; ws();
      if (i < s.size() && s[i] == '}') { i++; return j; }
      for (;;) {
        ws(); if (i >= s.size() || s[i] != '"') fail("key must be a string");
        size_t e = s.find('"', i + 1); if (e == std::string::npos) fail("EOF while parsing a string");
        std::string key = s.substr(i + 1, e - i - 1); i = e + 1; ws();
        if (i >= s.size() || s[i] != ':') fail("expected `:`");
        i++;
        Json v = value();
        j.fields.emplace_back(std::move(key), std::move(v)); ws();
        if (i < s.size() && s[i] == ',') { i++; continue; }
        if (i < s.size() && s[i] == '}') { i++; break; }
        fail("expected `,` or `}`");
      }
    } else if (isdigit((unsigned char)c)) {
      j.kind = Json::Num;
      while (i < s.size() && isdigit((unsigned char)s[i])) { j.num = j.num * 10 + (uint64_t)(s[i] - '0'); i++; }
    } else if (s.compare(i, 4, "null") == 0) {
      i += 4;
    } else fail("expected value");
    return j;
  }
};
static Json parse_json_file(const char* path) {
  std::ifstream f(path, std::ios::binary);
  std::stringstream ss; ss << f.rdbuf();
  std::string src = ss.str();
  JsonParser p(src);
  Json j = p.value(); p.ws();
  if (p.i != src.size()) p.fail("trailing characters");
  return j;
}
static std::vector<uint8_t> json_bytes(const Json& j) {
  if (j.kind != Json::Arr) throw CliError("invalid type: expected Sequence of u8");
  std::vector<uint8_t> out;
  for (auto& e : j.items) {
    if (e.kind != Json::Num || e.num > 255) throw CliError("invalid value: expected u8");
    out.push_back((uint8_t)e.num);
  }
  return out;
}
static void put_bytes(std::string& out, const uint8_t* b, size_t n) {
  out += '[';
  for (size_t i = 0; i < n; i++) { if (i) out += ','; out += std::to_string((unsigned)b[i]); }
  out += ']';
}

// ---- serde shapes ------------------------------------------------------------------------------------------
static Scalar scalar_from_json(const Json& j) {              // src/scalar.rs:235-270
  auto b = json_bytes(j);
  if (b.size() != 32) throw CliError("Invalid byte array, expected length 32, got " + std::to_string(b.size()));
  std::array<uint8_t, 32> le; std::copy(b.begin(), b.end(), le.begin());
  return Scalar::from_le_bytes(le);
}
static G1Point g1_from_json(const Json& j) { auto b = json_bytes(j); return G1Point::from_compressed(b.data(), b.size()); }
static G2Point g2_from_json(const Json& j) { auto b = json_bytes(j); return G2Point::from_compressed(b.data(), b.size()); }
static void put_scalar(std::string& o, const Scalar& s) { auto le = s.to_le_bytes(); put_bytes(o, le.data(), 32); }
static void put_polynomial(std::string& o, const Polynomial& p) {
  o += "{\"coefficients\":[";
  for (size_t i = 0; i < p.coefficients().size(); i++) { if (i) o += ','; put_scalar(o, p.coefficients()[i]); }
  o += "]}";
}
static Polynomial polynomial_from_json(const Json& j) {
  std::vector<Scalar> co;
  const Json& arr = j.at("coefficients");
  if (arr.kind != Json::Arr) throw CliError("invalid type: expected a sequence");
  for (auto& e : arr.items) co.push_back(scalar_from_json(e));
  return Polynomial::from_raw_coefficients(std::move(co));   // derived Deserialize: no re-trimming
}

static bool exists(const char* p) { struct stat st; return stat(p, &st) == 0; }
static void write_file(const char* path, const std::string& data) {
  if (exists(path)) unlink(path);
  std::ofstream f(path, std::ios::binary | std::ios::trunc);
  if (!f) throw CliError(std::string("No such file or directory (os error 2): ") + path);
  f << data;
}

static std::unique_ptr<TrustedSetup> read_setup(const std::shared_ptr<Context>& ctx) {
  if (!exists(SETUP_ARTIFACTS_PATH)) throw CliError("Trusted setup artifacts have not been found, generate them beforehand.");
  Json j = parse_json_file(SETUP_ARTIFACTS_PATH);
  if (j.kind != Json::Arr) throw CliError("invalid type: expected a sequence");
  std::vector<SetupArtifact> items;
  for (auto& e : j.items) {
    SetupArtifact a; a.g1 = g1_from_json(e.at("g1")); a.g2 = g2_from_json(e.at("g2")); a.has_g2 = true;
    items.push_back(a);
  }
  return TrustedSetup::from_artifacts(ctx, items);
}
struct CommitmentArtifact { Polynomial polynomial; G1Point commitment; };
static CommitmentArtifact read_commitment() {
  if (!exists(COMMITMENT_ARTIFACTS_PATH)) throw CliError("Commitment artifact has not been found, generate it beforehand.");
  Json j = parse_json_file(COMMITMENT_ARTIFACTS_PATH);
  return CommitmentArtifact{polynomial_from_json(j.at("polynomial")), g1_from_json(j.at("commitment"))};
}

static std::string i128_to_string(__int128 v) {
  if (v == 0) return "0";
  bool neg = v < 0; unsigned __int128 m = neg ? (unsigned __int128)(-(v + 1)) + 1 : (unsigned __int128)v;
  std::string s; while (m) { s.push_back((char)('0' + (int)(m % 10))); m /= 10; }
  if (neg) s.push_back('-');
  return std::string(s.rbegin(), s.rend());
}
static bool parse_i128(const char* a, __int128* out) {
  const char* p = a; bool neg = false;
  if (*p == '-') { neg = true; p++; } else if (*p == '+') p++;
  if (!*p) return false;
  unsigned __int128 m = 0; const unsigned __int128 lim = (unsigned __int128)1 << 127;
  for (; *p; p++) {
    if (!isdigit((unsigned char)*p)) return false;
    unsigned d = (unsigned)(*p - '0');
    if (m > (lim - d) / 10) return false;
    m = m * 10 + d;
  }
  if (!neg && m >= lim) return false;
  *out = neg ? -(__int128)(m - 1) - 1 : (__int128)m;
  return true;
}

// ---- commands ----------------------------------------------------------------------------------------------
static void cmd_trusted_setup() {
  info("Starting the trusted setup ceremony");
  if (!exists(ARTIFACTS_FOLDER_PATH) && mkdir(ARTIFACTS_FOLDER_PATH, 0777) != 0) throw CliError("cannot create ./artifacts");
  std::array<uint8_t, 32> secret{};
  if (const char* hexs = getenv("KZG_CLI_SECRET_HEX")) {
    if (strlen(hexs) != 64) throw CliError("KZG_CLI_SECRET_HEX must hold 64 hex digits");
    for (int i = 0; i < 32; i++) { unsigned v; if (sscanf(hexs + 2 * i, "%2x", &v) != 1) throw CliError("bad hex"); secret[i] = (uint8_t)v; }
  } else {
    std::ifstream r("/dev/urandom", std::ios::binary);
    r.read(reinterpret_cast<char*>(secret.data()), 32);
    if (!r) throw CliError("cannot read /dev/urandom");
  }
  auto ctx = std::make_shared<Context>(0);
  auto setup = trusted_setup::SetupArtifactsGenerator(ctx, secret).take(MAX_DEGREE + 1);
  std::string out = "[";
  for (size_t i = 0; i < setup->len(); i++) {
    SetupArtifact a = setup->at(i);
    if (i) out += ',';
    out += "{\"g1\":"; put_bytes(out, a.g1.compressed().data(), 48);
    out += ",\"g2\":"; put_bytes(out, a.g2.compressed().data(), 96);
    out += '}';
  }
  out += ']';
  write_file(SETUP_ARTIFACTS_PATH, out);
  info(std::string("Trusted setup ceremony successfully performed. Artifacts have been written in \"") + SETUP_ARTIFACTS_PATH + "\"");
}

static void cmd_commit(const std::vector<__int128>& coefficients) {
  Polynomial polynomial = Polynomial::try_from_i128(coefficients);
  std::string shown = polynomial.to_string();
  if (polynomial.degree() > MAX_DEGREE) throw CliError("Only polynomials up to degree " + std::to_string(MAX_DEGREE) + " are supported");
  info("Starting to commit to the polynomial P(x) = \"" + shown + "\"");
  if (!exists(SETUP_ARTIFACTS_PATH)) throw CliError("Trusted setup artifacts have not been found, generate them beforehand.");
  auto ctx = std::make_shared<Context>(0);
  auto setup = read_setup(ctx);
  G1Point commitment = polynomial.commit(*setup);
  std::string out = "{\"polynomial\":"; put_polynomial(out, polynomial);
  out += ",\"commitment\":"; put_bytes(out, commitment.compressed().data(), 48); out += '}';
  write_file(COMMITMENT_ARTIFACTS_PATH, out);
  info("Commitment to the polynomial \"P(x) = " + shown + "\" has been successfully generated.");
}

static void cmd_evaluate(__int128 x) {
  std::string xs = i128_to_string(x);
  info("Starting to evaluate the committed polynomial at input point \"x = " + xs + "\"");
  if (!exists(SETUP_ARTIFACTS_PATH)) throw CliError("Trusted setup artifacts have not been found, generate them beforehand.");
  auto ctx = std::make_shared<Context>(0);
  auto setup = read_setup(ctx);
  CommitmentArtifact ca = read_commitment();
  Evaluation evaluation = ca.polynomial.evaluate(*ctx, Scalar::from_i128(x));
  G1Point proof = evaluation.generate_proof(ca.polynomial, *setup);
  std::string out = "{\"evaluation\":{\"point\":"; put_scalar(out, evaluation.point);
  out += ",\"result\":"; put_scalar(out, evaluation.result);
  out += "},\"proof\":"; put_bytes(out, proof.compressed().data(), 48); out += '}';
  write_file(EVALUATION_ARTIFACTS_PATH, out);
  info("Evaluation successful for polynomial: \"P(x) = " + ca.polynomial.to_string() + "\" at point \"x = " + xs +
       "\" with \"P(" + xs + ") = " + evaluation.result.to_string() + "\"");
}

static void cmd_verify() {
  info("Starting to verify the previous polynomial evaluation");
  if (!exists(SETUP_ARTIFACTS_PATH)) throw CliError("Trusted setup artifacts have not been found, generate them beforehand.");
  auto ctx = std::make_shared<Context>(0);
  auto setup = read_setup(ctx);
  CommitmentArtifact ca = read_commitment();
  if (!exists(EVALUATION_ARTIFACTS_PATH)) throw CliError("Evaluation artifact has not been found, generate it beforehand.");
  Json j = parse_json_file(EVALUATION_ARTIFACTS_PATH);
  Evaluation ev{scalar_from_json(j.at("evaluation").at("point")), scalar_from_json(j.at("evaluation").at("result"))};
  G1Point proof = g1_from_json(j.at("proof"));
  if (!ev.verify_proof(proof, ca.commitment, *setup)) throw CliError("The proof associated to the evaluation is incorrect.");
  info("Successfully verified evaluation for polynomial \"P(x) = " + ca.polynomial.to_string() + "\" at point \"x = " +
       ev.point.to_string() + "\" with \"P(" + ev.point.to_string() + ") = " + ev.result.to_string() + "\"");
}

static void usage(FILE* f) {
  fprintf(f,
          "Usage: kzg-cli [OPTIONS] [COMMAND]\n\n"
          "Commands:\n"
          "  trusted-setup      Perform a trusted setup ceremony and write the artifacts in './artifacts/setup.json'\n"
          "  commit             Commit to a polynomial using the trusted setup artifacts\n"
          "  evaluate           Evaluate the committed polynomial at an input point and generate the associated Kate proof\n"
          "  verify-evaluation  Verify the previous evaluation with its proof\n"
          "  help               Print this message or the help of the given subcommand(s)\n\n"
          "Options:\n"
          "  -d, --debug...  Turn debugging information on\n"
          "  -h, --help      Print help\n"
          "  -V, --version   Print version\n");
}

int main(int argc, char** argv) {
  int debug = 0;
  std::vector<std::string> rest;
  bool after_dashes = false;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    if (!after_dashes && a == "--") { after_dashes = true; continue; }
    if (!after_dashes && rest.empty() && (a == "-d" || a == "--debug")) { debug++; continue; }
    if (!after_dashes && rest.empty() && a.size() > 1 && a[0] == '-' && a.find_first_not_of('d', 1) == std::string::npos) { debug += (int)a.size() - 1; continue; }
    if (!after_dashes && (a == "-h" || a == "--help")) { usage(stdout); return 0; }
    if (!after_dashes && rest.empty() && (a == "-V" || a == "--version")) { printf("kzg_poly_commit_exploration 0.1.0\n"); return 0; }
    rest.push_back(a);
  }
  g_level = debug == 0 ? 3 : debug == 1 ? 4 : 5;
  if (const char* lv = getenv("LOG_LEVEL")) {
    std::string v = lv; for (auto& c : v) c = (char)tolower((unsigned char)c);
    if (v == "error") g_level = 1; else if (v == "warn") g_level = 2; else if (v == "info") g_level = 3;
    else if (v == "debug") g_level = 4; else if (v == "trace") g_level = 5;
  }
  if (rest.empty()) { warn("No command has been input"); return 0; }
  try {
    const std::string& cmd = rest[0];
    if (cmd == "help") { usage(stdout); return 0; }
    if (cmd == "trusted-setup") {
      if (rest.size() != 1) { fprintf(stderr, "error: unexpected argument '%s' found\n", rest[1].c_str()); return 2; }
      cmd_trusted_setup();
    } else if (cmd == "commit") {
      if (rest.size() < 2) { fprintf(stderr, "error: the following required arguments were not provided:\n  <COEFFICIENTS>...\n"); return 2; }
      std::vector<__int128> co;
      for (size_t i = 1; i < rest.size(); i++) {
        __int128 v;
        if (!parse_i128(rest[i].c_str(), &v)) { fprintf(stderr, "error: invalid value '%s' for '[COEFFICIENTS]...': invalid digit found in string\n", rest[i].c_str()); return 2; }
        co.push_back(v);
      }
      cmd_commit(co);
    } else if (cmd == "evaluate") {
      __int128 x;
      if (rest.size() != 2 || !parse_i128(rest[1].c_str(), &x)) { fprintf(stderr, "error: invalid or missing value for '<X>'\n"); return 2; }
      cmd_evaluate(x);
    } else if (cmd == "verify-evaluation") {
      if (rest.size() != 1) { fprintf(stderr, "error: unexpected argument '%s' found\n", rest[1].c_str()); return 2; }
      cmd_verify();
    } else {
      fprintf(stderr, "error: unrecognized subcommand '%s'\n\n", cmd.c_str()); usage(stderr); return 2;
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "Command execution failed with error: Unhandled error: %s\n", e.what());
    return 101;
  }
  return 0;
}
