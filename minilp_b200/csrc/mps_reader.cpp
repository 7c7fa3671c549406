// Free-format MPS reader: host-side restatement of the reference's MpsFile::parse
// (src/mps.rs:39-329) over one in-memory buffer, producing the arrays Solver::try_new / mlp_solver_create_sparse take —
// no per-entry host objects, one pass, names hashed as views into the text.  Behaviour follows the reference: whitespace
// tokens, `*` comment lines and blank lines skipped with 1-based line numbers kept (332-358), sections NAME / ROWS /
// COLUMNS / RHS / [RANGES] / [BOUNDS] / ENDATA, a data line starts with a space, the first N row is the objective and
// later N rows are ignored free rows, only the FIRST RHS / RANGES / BOUNDS vector counts (193-198, 223-228, 253-258), bound
// types LO / UP / FX / FR (282), a negative UP bound without LO gives (-inf, max] (299), a ranged row becomes two
// constraints (306-321).  Syntax errors return MLP_INVALID with the reference's "line N: ..." text in mlp_last_error().
#include "minilp_b200.h"

#include <charconv>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <algorithm>
#include <cstdlib>
#include <string_view>
#include <thread>
#include <vector>

extern "C" void mlp_set_last_error(const char* msg);

namespace {
constexpr double kInf = std::numeric_limits<double>::infinity();
using sv = std::string_view;

struct ParseError {
  std::string msg;
};

struct Lines {  // mps.rs:332-358
  const char* p;
  const char* end;
  sv cur;
  int64_t idx = 0;
  Lines(const char* b, int64_t n) : p(b), end(b + n) {}
  void to_next() {
    for (;;) {
      ++idx;
      if (p >= end) { cur = sv(); return; }
      const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
      const char* le = nl ? nl : end;
      sv line(p, (size_t)(le - p));
      p = nl ? nl + 1 : end;
      if (!line.empty() && line[0] == '*') continue;
      size_t e = line.size();
      while (e > 0 && (line[e - 1] == ' ' || (line[e - 1] >= '\t' && line[e - 1] <= '\r'))) --e;  // trim_end
      if (e == 0) continue;
      cur = line.substr(0, e);
      return;
    }
  }
  [[noreturn]] void err(const std::string& m) const { throw ParseError{"line " + std::to_string(idx) + ": " + m}; }
};

struct Tokens {  // split_whitespace over the current line
  sv s;
  size_t pos = 0;
  int64_t line_idx;
  explicit Tokens(const Lines& l) : s(l.cur), line_idx(l.idx) {}
  static bool ws(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }
  bool opt(sv& out) {
    while (pos < s.size() && ws(s[pos])) ++pos;
    if (pos >= s.size()) return false;
    const size_t b = pos;
    while (pos < s.size() && !ws(s[pos])) ++pos;
    out = s.substr(b, pos - b);
    return true;
  }
  sv next() {
    sv t;
    if (!opt(t)) throw ParseError{"line " + std::to_string(line_idx) + ": unexpected end of line"};
    return t;
  }
};

double parse_f64(sv tok, int64_t line_idx) {  // f64::from_str (mps.rs:395-402): whole token, optional sign, inf / nan
  const char* b = tok.data();
  const char* e = b + tok.size();
  double v = 0.0;
  bool ok = b < e;
  if (ok) {
    if (*b == '+') ++b;  // from_chars takes no '+', Rust does
    ok = b < e && *b != '+' && !(*b == '-' && b + 1 < e && b[1] == '+');
    if (ok) {
      auto r = std::from_chars(b, e, v);
      ok = r.ec == std::errc() && r.ptr == e;
      if (r.ec == std::errc::result_out_of_range && r.ptr == e) {  // Rust saturates to +-inf / 0 instead of failing
        ok = true;
        bool neg = *b == '-';
        bool tiny = false;
        for (const char* q = b; q < e; ++q)
          if ((*q == 'e' || *q == 'E') && q + 1 < e && q[1] == '-') tiny = true;
        v = tiny ? 0.0 : kInf;
        if (neg) v = -v;
      }
    }
  }
  if (!ok) throw ParseError{"line " + std::to_string(line_idx) + ": couldn't parse float from string: `" + std::string(tok) + "`"};
  return v;
}

// Name table: open addressing over views into the text (the COLUMNS section does one row-name lookup per entry — 10^7 for
// BASELINE config 4 — and a node-based map spends most of the parse in cache misses).
struct NameTable {
  struct Slot { uint64_t h; const char* p; uint32_t len; int32_t val; };
  std::vector<Slot> slots;
  size_t mask = 0, count = 0;
  NameTable() { rehash(1024); }
  static uint64_t hash(sv s) {  // FNV-1a over 8-byte words, finalised with a multiply-shift mix
    uint64_t h = 0xcbf29ce484222325ull ^ s.size();
    const char* p = s.data();
    size_t n = s.size();
    while (n >= 8) { uint64_t w; std::memcpy(&w, p, 8); h = (h ^ w) * 0x100000001b3ull; p += 8; n -= 8; }
    uint64_t w = 0;
    std::memcpy(&w, p, n);
    h = (h ^ w) * 0x100000001b3ull;
    h ^= h >> 32;
    h *= 0x9e3779b97f4a7c15ull;
    return h ^ (h >> 29);
  }
  void rehash(size_t cap) {
    std::vector<Slot> old;
    old.swap(slots);
    slots.assign(cap, Slot{0, nullptr, 0, 0});
    mask = cap - 1;
    for (const Slot& s : old)
      if (s.p) {
        size_t i = s.h & mask;
        while (slots[i].p) i = (i + 1) & mask;
        slots[i] = s;
      }
  }
  int32_t* find(sv k) {
    const uint64_t h = hash(k);
    for (size_t i = h & mask;; i = (i + 1) & mask) {
      Slot& s = slots[i];
      if (!s.p) return nullptr;
      if (s.h == h && s.len == k.size() && std::memcmp(s.p, k.data(), k.size()) == 0) return &s.val;
    }
  }
  bool insert(sv k, int32_t v) {  // false: already present
    if ((count + 1) * 2 > slots.size()) rehash(slots.size() * 2);
    const uint64_t h = hash(k);
    for (size_t i = h & mask;; i = (i + 1) & mask) {
      Slot& s = slots[i];
      if (!s.p) { s = Slot{h, k.data(), (uint32_t)k.size(), v}; ++count; return true; }
      if (s.h == h && s.len == k.size() && std::memcmp(s.p, k.data(), k.size()) == 0) return false;
    }
  }
};

struct KV {
  sv k[2];
  double v[2];
  int n;
};
KV kv_pairs(Tokens& t) {  // mps.rs:404-431
  KV r;
  r.k[0] = t.next();
  r.v[0] = parse_f64(t.next(), t.line_idx);
  r.n = 1;
  sv k2;
  if (t.opt(k2)) {
    r.k[1] = k2;
    r.v[1] = parse_f64(t.next(), t.line_idx);
    r.n = 2;
  }
  return r;
}
}  // namespace

struct mlp_mps {
  std::string name;
  std::vector<char> names_blob;        // variable names, concatenated
  std::vector<int64_t> names_off;      // n + 1
  std::vector<double> obj, mins, maxs;  // n (objective in the USER's sign; bounds as Problem::add_var receives them)
  std::vector<int64_t> row_ptr;         // rows + 1 (every constraint, empty ones included; ranged rows already doubled)
  std::vector<int32_t> col_idx;
  std::vector<double> vals;
  std::vector<int32_t> ops;             // 0 Eq, 1 Le, 2 Ge
  std::vector<double> rhs;
};

static void parse_into(const char* text, int64_t len, mlp_mps& out) {
  Lines lines(text, len);
  lines.to_next();
  {
    Tokens tk(lines);
    if (tk.next() != "NAME") lines.err("expected NAME section");
    sv nm;
    out.name = tk.opt(nm) ? std::string(nm) : std::string();
  }
  sv obj_name;
  bool have_obj = false;
  NameTable free_rows;
  struct Row { int op; double rhs, range; };
  std::vector<Row> rows;
  NameTable cidx;
  lines.to_next();
  if (lines.cur != "ROWS") lines.err("expected ROWS section");
  for (;;) {
    lines.to_next();
    if (lines.cur.empty() || lines.cur[0] != ' ') break;
    Tokens tk(lines);
    const sv row_type = tk.next(), name = tk.next();
    if (row_type == "N") {
      if (!have_obj) { obj_name = name; have_obj = true; }
      else free_rows.insert(name, 1);
      continue;
    }
    int op;
    if (row_type == "L") op = 1;
    else if (row_type == "G") op = 2;
    else if (row_type == "E") op = 0;
    else lines.err("unexpected row type " + std::string(row_type));
    if (!cidx.insert(name, (int32_t)rows.size())) lines.err("row " + std::string(name) + " already declared");
    rows.push_back(Row{op, 0.0, 0.0});
  }
  if (!have_obj) lines.err("objective function name not declared");

  // COLUMNS: entries arrive grouped by variable, so each row's entries arrive in ascending variable order
  struct VarDef { double mn, mx, obj; bool has_mn, has_mx; };
  std::vector<VarDef> vars;
  std::vector<sv> var_names;
  NameTable vidx;
  std::vector<int32_t> e_row, e_var;
  std::vector<double> e_val;
  if (lines.cur != "COLUMNS") lines.err("expected COLUMNS section");
  // Fast path for large files: the COLUMNS body (99 % of a big file) is cut at line boundaries into one piece per host
  // thread; every piece is tokenised independently (the row-name table is read-only by now) into entries with piece-local
  // variable numbers, and a serial stitch restores the global numbering — a variable whose lines straddle a cut continues
  // across it.  Anything unusual (a syntax error, a re-declared variable, a non-data line inside the body) abandons the fast
  // path and the serial loop below re-reads the section, so error messages and line numbers are exactly the serial ones.
  bool columns_done = false;
  {
    const char* body = lines.p;
    const char* stop = nullptr;  // start of the first "RHS" header line after the body
    for (const char* q = body; q < lines.end;) {
      const char* hit = (const char*)memmem(q, (size_t)(lines.end - q), "\nRHS", 4);
      if (!hit) break;
      const char* a = hit + 4;
      while (a < lines.end && *a != '\n' && Tokens::ws(*a)) ++a;
      if (a >= lines.end || *a == '\n') { stop = hit + 1; break; }
      q = hit + 1;
    }
    unsigned T = std::min<unsigned>(std::min<unsigned>(std::thread::hardware_concurrency(), 32u),
                                    stop ? (unsigned)((stop - body) >> 22) : 0u);  // >= 4 MB per piece
    if (const char* env = std::getenv("MLP_MPS_THREADS")) T = std::min<unsigned>(T ? T : 1u, (unsigned)std::max(1, atoi(env)));
    if (stop && T >= 2) {
      struct Piece {
        std::vector<int32_t> e_row, e_loc;
        std::vector<double> e_val, objs;
        std::vector<char> obj_set;
        std::vector<sv> names;
        int64_t lines = 0;
        bool fail = false;
      };
      std::vector<Piece> pieces(T);
      std::vector<const char*> cut(T + 1);
      cut[0] = body;
      cut[T] = stop;
      for (unsigned t = 1; t < T; ++t) {
        const char* q = body + (size_t)(stop - body) * t / T;
        const char* nl = (const char*)memchr(q, '\n', (size_t)(stop - q));
        cut[t] = nl ? nl + 1 : stop;
      }
      auto work = [&](unsigned t) {
        Piece& P = pieces[t];
        try {
          Lines L(cut[t], cut[t + 1] - cut[t]);
          const size_t guess = (size_t)(cut[t + 1] - cut[t]) / 18;
          P.e_row.reserve(guess); P.e_loc.reserve(guess); P.e_val.reserve(guess);
          sv cur;
          bool have = false;
          for (;;) {
            L.to_next();
            if (L.cur.empty()) break;
            if (L.cur[0] != ' ') { P.fail = true; return; }
            Tokens tk(L);
            const sv name = tk.next();
            if (!have || name != cur) { P.names.push_back(name); P.objs.push_back(0.0); P.obj_set.push_back(0); cur = name; have = true; }
            const KV kv = kv_pairs(tk);
            for (int q = 0; q < kv.n; ++q) {
              if (kv.k[q] == obj_name) { P.objs.back() = kv.v[q]; P.obj_set.back() = 1; }
              else {
                const int32_t* it = cidx.find(kv.k[q]);
                if (it) { P.e_row.push_back(*it); P.e_loc.push_back((int32_t)P.names.size() - 1); P.e_val.push_back(kv.v[q]); }
                else if (!free_rows.find(kv.k[q])) { P.fail = true; return; }
              }
            }
          }
          P.lines = L.idx - 1;
        } catch (const ParseError&) {
          P.fail = true;
        }
      };
      std::vector<std::thread> th;
      for (unsigned t = 1; t < T; ++t) th.emplace_back(work, t);
      work(0);
      for (auto& x : th) x.join();
      bool ok = true;
      for (const Piece& P : pieces) ok = ok && !P.fail;
      // stitch: global variable numbers; a piece whose first name equals the previous variable continues it
      std::vector<int32_t> base(T, 0);
      if (ok) {
        for (unsigned t = 0; t < T && ok; ++t) {
          Piece& P = pieces[t];
          for (size_t j = 0; j < P.names.size() && ok; ++j) {
            const bool cont = j == 0 && !var_names.empty() && var_names.back() == P.names[0];
            if (j == 0) base[t] = (int32_t)var_names.size() - (cont ? 1 : 0);
            if (cont) { if (P.obj_set[0]) vars.back().obj = P.objs[0]; continue; }
            if (!vidx.insert(P.names[j], (int32_t)var_names.size())) { ok = false; break; }  // re-declared: let the serial loop say so
            var_names.push_back(P.names[j]);
            vars.push_back(VarDef{0, 0, P.objs[j], false, false});
          }
        }
      }
      if (ok) {
        size_t total = 0;
        for (const Piece& P : pieces) total += P.e_row.size();
        e_row.resize(total); e_var.resize(total); e_val.resize(total);
        std::vector<size_t> off(T + 1, 0);
        for (unsigned t = 0; t < T; ++t) off[t + 1] = off[t] + pieces[t].e_row.size();
        auto copy_out = [&](unsigned t) {
          const Piece& P = pieces[t];
          const size_t o = off[t], cnt = P.e_row.size();
          if (cnt) { std::memcpy(e_row.data() + o, P.e_row.data(), cnt * 4); std::memcpy(e_val.data() + o, P.e_val.data(), cnt * 8); }
          for (size_t q = 0; q < cnt; ++q) e_var[o + q] = base[t] + P.e_loc[q];
        };
        std::vector<std::thread> th2;
        for (unsigned t = 1; t < T; ++t) th2.emplace_back(copy_out, t);
        copy_out(0);
        for (auto& x : th2) x.join();
        for (const Piece& P : pieces) lines.idx += P.lines;
        lines.p = stop;
        lines.to_next();  // the RHS header
        columns_done = true;
      } else {  // undo the stitch; the serial loop starts from the section header again
        vars.clear();
        var_names.clear();
        vidx = NameTable();
      }
    }
  }
  sv cur_name;
  bool have_cur = false;
  VarDef cur_def{0, 0, 0.0, false, false};
  int32_t cur_var = 0;
  for (; !columns_done;) {
    lines.to_next();
    if (lines.cur.empty() || lines.cur[0] != ' ') break;
    Tokens tk(lines);
    const sv name = tk.next();
    if (!have_cur || name != cur_name) {
      if (vidx.find(name)) lines.err("variable " + std::string(name) + " already declared");
      if (have_cur) {
        vidx.insert(cur_name, cur_var);
        var_names.push_back(cur_name);
        vars.push_back(cur_def);
        cur_def = VarDef{0, 0, 0.0, false, false};
        ++cur_var;
      }
      cur_name = name;
      have_cur = true;
    }
    const KV kv = kv_pairs(tk);
    for (int q = 0; q < kv.n; ++q) {
      if (kv.k[q] == obj_name) cur_def.obj = kv.v[q];
      else {
        const int32_t* it = cidx.find(kv.k[q]);
        if (it) { e_row.push_back(*it); e_var.push_back(cur_var); e_val.push_back(kv.v[q]); }
        else if (!free_rows.find(kv.k[q])) lines.err("unknown constraint: " + std::string(kv.k[q]));
      }
    }
  }
  if (have_cur) {
    vidx.insert(cur_name, cur_var);
    var_names.push_back(cur_name);
    vars.push_back(cur_def);
  }

  auto first_vector_section = [&](auto&& on_pair, bool bounds) {
    sv vec;
    bool have_vec = false;
    for (;;) {
      lines.to_next();
      if (lines.cur.empty() || lines.cur[0] != ' ') break;
      Tokens tk(lines);
      sv btype;
      if (bounds) btype = tk.next();
      const sv vn = tk.next();
      if (!have_vec) { vec = vn; have_vec = true; }
      else if (vec != vn) continue;
      on_pair(tk, btype);
    }
  };
  if (lines.cur != "RHS") lines.err("expected RHS section");
  first_vector_section([&](Tokens& tk, sv) {
    const KV kv = kv_pairs(tk);
    for (int q = 0; q < kv.n; ++q) {
      if (kv.k[q] == obj_name) lines.err("setting objective in RHS section is not supported");
      const int32_t* it = cidx.find(kv.k[q]);
      if (!it) lines.err("unknown constraint: " + std::string(kv.k[q]));
      rows[(size_t)*it].rhs = kv.v[q];
    }
  }, false);
  if (lines.cur == "RANGES")
    first_vector_section([&](Tokens& tk, sv) {
      const KV kv = kv_pairs(tk);
      for (int q = 0; q < kv.n; ++q) {
        const int32_t* it = cidx.find(kv.k[q]);
        if (!it) lines.err("unknown constraint: " + std::string(kv.k[q]));
        rows[(size_t)*it].range = kv.v[q];
      }
    }, false);
  if (lines.cur == "BOUNDS")
    first_vector_section([&](Tokens& tk, sv btype) {
      const sv vname = tk.next();
      const int32_t* it = vidx.find(vname);
      if (!it) lines.err("unknown variable: " + std::string(vname));
      VarDef& vd = vars[(size_t)*it];
      if (btype == "FR") { vd.mn = -kInf; vd.mx = kInf; vd.has_mn = vd.has_mx = true; return; }
      const double val = parse_f64(tk.next(), lines.idx);
      if (btype == "LO") { vd.mn = val; vd.has_mn = true; }
      else if (btype == "UP") { vd.mx = val; vd.has_mx = true; }
      else if (btype == "FX") { vd.mn = vd.mx = val; vd.has_mn = vd.has_mx = true; }
      else lines.err("bound type " + std::string(btype) + " is not supported");
    }, true);
  if (lines.cur != "ENDATA") lines.err("expected ENDATA section");

  // ---- Problem (mps.rs:292-322)
  const size_t n = vars.size();
  out.obj.resize(n); out.mins.resize(n); out.maxs.resize(n);
  out.names_off.assign(n + 1, 0);
  for (size_t j = 0; j < n; ++j) {
    const VarDef& v = vars[j];
    out.obj[j] = v.obj;
    if (v.has_mn && v.has_mx) { out.mins[j] = v.mn; out.maxs[j] = v.mx; }
    else if (v.has_mn) { out.mins[j] = v.mn; out.maxs[j] = kInf; }
    else if (v.has_mx) { out.mins[j] = v.mx < 0.0 ? -kInf : 0.0; out.maxs[j] = v.mx; }
    else { out.mins[j] = 0.0; out.maxs[j] = kInf; }
    out.names_off[j + 1] = out.names_off[j] + (int64_t)var_names[j].size();
  }
  out.names_blob.resize((size_t)out.names_off[n]);
  for (size_t j = 0; j < n; ++j) std::memcpy(out.names_blob.data() + out.names_off[j], var_names[j].data(), var_names[j].size());
  // declared row -> first output row; a ranged row yields two (Ge lo, then Le hi)
  const size_t R = rows.size();
  std::vector<int64_t> cnt(R, 0), first_out(R + 1, 0), src_ptr(R + 1, 0);
  for (int32_t r : e_row) cnt[(size_t)r] += 1;
  for (size_t r = 0; r < R; ++r) {
    first_out[r + 1] = first_out[r] + (rows[r].range == 0.0 ? 1 : 2);
    src_ptr[r + 1] = src_ptr[r] + cnt[r];
  }
  // stable counting sort of the entries by declared row (keeps ascending variable order within a row)
  const size_t E = e_row.size();
  std::vector<int32_t> s_var(E);
  std::vector<double> s_val(E);
  {
    std::vector<int64_t> fill(src_ptr.begin(), src_ptr.end() - 1);
    for (size_t t = 0; t < E; ++t) {
      const int64_t d = fill[(size_t)e_row[t]]++;
      s_var[(size_t)d] = e_var[t];
      s_val[(size_t)d] = e_val[t];
    }
  }
  for (size_t r = 0; r < R; ++r)  // CsVec::new panics on a repeated index (lib.rs:247-249, 279)
    for (int64_t t = src_ptr[r] + 1; t < src_ptr[r + 1]; ++t)
      if (s_var[(size_t)t] == s_var[(size_t)t - 1]) throw ParseError{"variable added more than once to a constraint"};
  const size_t OUT = (size_t)first_out[R];
  out.row_ptr.assign(OUT + 1, 0);
  out.ops.resize(OUT);
  out.rhs.resize(OUT);
  for (size_t r = 0; r < R; ++r) {
    const Row& rw = rows[r];
    const size_t o = (size_t)first_out[r];
    if (rw.range == 0.0) {
      out.ops[o] = rw.op; out.rhs[o] = rw.rhs; out.row_ptr[o + 1] = cnt[r];
    } else {
      double lo, hi;  // mps.rs:306-316
      if (rw.op == 2) { lo = rw.rhs; hi = rw.rhs + std::fabs(rw.range); }
      else if (rw.op == 1) { lo = rw.rhs - std::fabs(rw.range); hi = rw.rhs; }
      else if (rw.range > 0.0) { lo = rw.rhs; hi = rw.rhs + rw.range; }
      else { lo = rw.rhs + rw.range; hi = rw.rhs; }
      out.ops[o] = 2; out.rhs[o] = lo; out.row_ptr[o + 1] = cnt[r];
      out.ops[o + 1] = 1; out.rhs[o + 1] = hi; out.row_ptr[o + 2] = cnt[r];
    }
  }
  for (size_t o = 0; o < OUT; ++o) out.row_ptr[o + 1] += out.row_ptr[o];
  out.col_idx.resize((size_t)out.row_ptr[OUT]);
  out.vals.resize((size_t)out.row_ptr[OUT]);
  for (size_t r = 0; r < R; ++r) {
    const int copies = rows[r].range == 0.0 ? 1 : 2;
    for (int c = 0; c < copies; ++c) {
      const int64_t dst = out.row_ptr[(size_t)first_out[r] + (size_t)c];
      std::memcpy(out.col_idx.data() + dst, s_var.data() + src_ptr[r], (size_t)cnt[r] * sizeof(int32_t));
      std::memcpy(out.vals.data() + dst, s_val.data() + src_ptr[r], (size_t)cnt[r] * sizeof(double));
    }
  }
}

extern "C" {

mlp_status mlp_mps_parse(const char* text, int64_t len, mlp_mps** out) {
  if (!out) return MLP_INVALID;
  *out = nullptr;
  if (!text || len < 0) { mlp_set_last_error("mps: no input"); return MLP_INVALID; }
  mlp_mps* f = new mlp_mps();
  try {
    parse_into(text, len, *f);
  } catch (const ParseError& e) {
    mlp_set_last_error(e.msg.c_str());
    delete f;
    return MLP_INVALID;
  } catch (const std::bad_alloc&) {
    mlp_set_last_error("mps: out of memory");
    delete f;
    return MLP_NOMEM;
  }
  *out = f;
  return MLP_OK;
}
void mlp_mps_free(mlp_mps* f) { delete f; }
const char* mlp_mps_name(mlp_mps* f) { return f ? f->name.c_str() : ""; }
void mlp_mps_sizes(mlp_mps* f, int64_t* num_vars, int64_t* num_constraints, int64_t* nnz, int64_t* names_bytes) {
  *num_vars = (int64_t)f->obj.size();
  *num_constraints = (int64_t)f->ops.size();
  *nnz = (int64_t)f->vals.size();
  *names_bytes = (int64_t)f->names_blob.size();
}
mlp_status mlp_mps_export(mlp_mps* f, double* obj, double* mins, double* maxs, int64_t* row_ptr, int32_t* col_idx, double* vals,
                          int32_t* ops, double* rhs, char* names_blob, int64_t* names_off) {
  if (!f) return MLP_INVALID;
  auto cp = [](void* d, const void* s, size_t b) { if (d && b) std::memcpy(d, s, b); };
  cp(obj, f->obj.data(), f->obj.size() * 8); cp(mins, f->mins.data(), f->mins.size() * 8); cp(maxs, f->maxs.data(), f->maxs.size() * 8);
  cp(row_ptr, f->row_ptr.data(), f->row_ptr.size() * 8); cp(col_idx, f->col_idx.data(), f->col_idx.size() * 4);
  cp(vals, f->vals.data(), f->vals.size() * 8); cp(ops, f->ops.data(), f->ops.size() * 4); cp(rhs, f->rhs.data(), f->rhs.size() * 8);
  cp(names_blob, f->names_blob.data(), f->names_blob.size()); cp(names_off, f->names_off.data(), f->names_off.size() * 8);
  return MLP_OK;
}

}  // extern "C"
