// checkpoint.cu -- safetensors GPTQ checkpoint ingest (host C++).
//
// What it stands in for: the reference's NativeSafetensorsLoader
// (crates/ferrum-quantization/src/native_safetensors.rs:130-365 open / shard index / typed reads /
// read_gptq_raw, :388-585 load_stacked_gptq_experts, :597-640 load_linear's GPTQ branch and fusion
// shims, :763-880 load_gptq_linear, :887-1102 load_gptq_linear_fused, :1242-1246 sym
// canonicalisation, :1288-1325 validate_gptq_g_idx, :1475-1530 load_quantize_config).  The
// reference parses the container with the `safetensors` and `serde_json` crates; neither is in
// the tree, so the published container format is restated here: 8-byte little-endian header
// length, a JSON object {name: {"dtype", "shape", "data_offsets": [begin, end]}, "__metadata__"?},
// then the data blob the offsets index.  Files are mmap'ed and tensors are copied out of the
// mapping on demand (the reference's Shard, :63-127).
//
// Everything up to the fb2_ckpt_load_* entry points is pure host code and is tested without a
// GPU (tests/test_checkpoint.py writes small checkpoints with an independent Python writer).
#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace fb2 {
namespace {

// ------------------------------------------------------------------------------ JSON
// A small recursive-descent parser: enough for safetensors headers, the shard index and the
// two config files.  Numbers keep both an integer and a double reading.
struct Json {
  enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
  bool b = false;
  double num = 0;
  long long inum = 0;
  bool is_int = false;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;     // insertion order kept

  const Json *get(const char *key) const {
    if (kind != Obj) return nullptr;
    for (auto &kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
};

struct JsonParser {
  const char *p, *end;
  std::string err;
  int depth = 0;

  bool fail(const char *what) {
    if (err.empty()) err = what;
    return false;
  }
  void ws() {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++;
  }
  static void put_utf8(std::string &s, unsigned cp) {
    if (cp < 0x80) s += (char)cp;
    else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 0x3F)); }
    else if (cp < 0x10000) { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
    else { s += (char)(0xF0 | (cp >> 18)); s += (char)(0x80 | ((cp >> 12) & 0x3F)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
  }
  bool hex4(unsigned &v) {
    if (end - p < 4) return fail("truncated \\u escape");
    v = 0;
    for (int i = 0; i < 4; i++) {
      const char c = *p++;
      v <<= 4;
      if (c >= '0' && c <= '9') v |= (unsigned)(c - '0');
      else if (c >= 'a' && c <= 'f') v |= (unsigned)(c - 'a' + 10);
      else if (c >= 'A' && c <= 'F') v |= (unsigned)(c - 'A' + 10);
      else return fail("bad \\u escape");
    }
    return true;
  }
  bool string(std::string &out) {
    if (p >= end || *p != '"') return fail("expected string");
    p++;
    out.clear();
    while (p < end && *p != '"') {
      if (*p == '\\') {
        if (++p >= end) return fail("truncated escape");
        const char c = *p++;
        switch (c) {
          case '"': out += '"'; break;
          case '\\': out += '\\'; break;
          case '/': out += '/'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'n': out += '\n'; break;
          case 'r': out += '\r'; break;
          case 't': out += '\t'; break;
          case 'u': {
            unsigned cp;
            if (!hex4(cp)) return false;
            if (cp >= 0xD800 && cp < 0xDC00) {          // high surrogate: a low one must follow (serde_json rejects lone ones)
              if (end - p < 6 || p[0] != '\\' || p[1] != 'u') return fail("lone surrogate in \\u escape");
              p += 2;
              unsigned lo;
              if (!hex4(lo)) return false;
              if (lo < 0xDC00 || lo > 0xDFFF) return fail("bad surrogate pair in \\u escape");
              cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
            } else if (cp >= 0xDC00 && cp <= 0xDFFF) {
              return fail("lone surrogate in \\u escape");
            }
            put_utf8(out, cp);
            break;
          }
          default: return fail("bad escape");
        }
      } else {
        if ((unsigned char)*p < 0x20) return fail("control character in string");
        out += *p++;
      }
    }
    if (p >= end) return fail("unterminated string");
    p++;
    if (!valid_utf8(out)) return fail("string is not valid UTF-8");
    return true;
  }
  // RFC 3629: no overlong forms, no surrogates, nothing above U+10FFFF (what serde_json enforces on the reference side)
  static bool valid_utf8(const std::string &s) {
    const unsigned char *q = (const unsigned char *)s.data(), *e = q + s.size();
    while (q < e) {
      const unsigned c = *q;
      int n;
      unsigned cp, lo;
      if (c < 0x80) { q++; continue; }
      else if ((c & 0xE0) == 0xC0) { n = 1; cp = c & 0x1F; lo = 0x80; }
      else if ((c & 0xF0) == 0xE0) { n = 2; cp = c & 0x0F; lo = 0x800; }
      else if ((c & 0xF8) == 0xF0) { n = 3; cp = c & 0x07; lo = 0x10000; }
      else return false;
      if (e - q <= n) return false;
      for (int i = 1; i <= n; i++) {
        if ((q[i] & 0xC0) != 0x80) return false;
        cp = (cp << 6) | (q[i] & 0x3F);
      }
      if (cp < lo || cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) return false;
      q += n + 1;
    }
    return true;
  }
  bool value(Json &v) {
    if (++depth > 64) return fail("nesting too deep");
    ws();
    if (p >= end) return fail("unexpected end of input");
    bool ok = true;
    const char c = *p;
    if (c == '{') {
      v.kind = Json::Obj;
      p++;
      ws();
      if (p < end && *p == '}') { p++; }
      else {
        for (;;) {
          ws();
          std::string key;
          if (!string(key)) { ok = false; break; }
          ws();
          if (p >= end || *p != ':') { ok = fail("expected ':'"); break; }
          p++;
          v.obj.emplace_back(std::move(key), Json());
          if (!value(v.obj.back().second)) { ok = false; break; }
          ws();
          if (p < end && *p == ',') { p++; continue; }
          if (p < end && *p == '}') { p++; break; }
          ok = fail("expected ',' or '}'");
          break;
        }
      }
    } else if (c == '[') {
      v.kind = Json::Arr;
      p++;
      ws();
      if (p < end && *p == ']') { p++; }
      else {
        for (;;) {
          v.arr.emplace_back();
          if (!value(v.arr.back())) { ok = false; break; }
          ws();
          if (p < end && *p == ',') { p++; continue; }
          if (p < end && *p == ']') { p++; break; }
          ok = fail("expected ',' or ']'");
          break;
        }
      }
    } else if (c == '"') {
      v.kind = Json::Str;
      ok = string(v.str);
    } else if (end - p >= 4 && !strncmp(p, "true", 4)) {
      v.kind = Json::Bool; v.b = true; p += 4;
    } else if (end - p >= 5 && !strncmp(p, "false", 5)) {
      v.kind = Json::Bool; v.b = false; p += 5;
    } else if (end - p >= 4 && !strncmp(p, "null", 4)) {
      v.kind = Json::Null; p += 4;
    } else if (c == '-' || (c >= '0' && c <= '9')) {
      const char *s = p;
      if (*p == '-') p++;
      bool integral = true;
      while (p < end && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '+' || *p == '-')) {
        if (*p == '.' || *p == 'e' || *p == 'E') integral = false;
        p++;
      }
      std::string t(s, p);
      v.kind = Json::Num;
      char *ep = nullptr;
      v.num = strtod(t.c_str(), &ep);
      if (!ep || *ep) ok = fail("bad number");
      if (ok && integral && t.size() <= 19) {
        v.is_int = true;
        v.inum = strtoll(t.c_str(), nullptr, 10);
      }
    } else {
      ok = fail("unexpected character");
    }
    depth--;
    return ok;
  }
};

bool parse_json(const char *data, size_t len, Json &out, std::string &err) {
  JsonParser jp{data, data + len, {}, 0};
  if (!jp.value(out)) { err = jp.err; return false; }
  jp.ws();
  if (jp.p != jp.end) { err = "trailing characters after the JSON value"; return false; }
  return true;
}

// ---------------------------------------------------------------------------- files
bool file_exists(const std::string &p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}

int read_text_file(const std::string &path, std::string &out) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) { set_error("read \"%s\": %s", path.c_str(), strerror(errno)); return FB2_ERR_IO; }
  char buf[1 << 14];
  size_t n;
  out.clear();
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) out.append(buf, n);
  fclose(f);
  return FB2_OK;
}

struct Mapping {
  const uint8_t *base = nullptr;
  size_t len = 0;
  ~Mapping() { if (base && len) munmap(const_cast<uint8_t *>(base), len); }
};

int dtype_code(const std::string &s, size_t *elem_bytes) {
  static const struct { const char *name; int code; size_t bytes; } kTable[] = {
      {"F32", FB2_DTYPE_F32, 4}, {"F16", FB2_DTYPE_F16, 2}, {"BF16", FB2_DTYPE_BF16, 2}, {"I32", FB2_DTYPE_I32, 4},
      {"F64", FB2_DTYPE_OTHER, 8}, {"I64", FB2_DTYPE_OTHER, 8}, {"U64", FB2_DTYPE_OTHER, 8}, {"U32", FB2_DTYPE_OTHER, 4},
      {"I16", FB2_DTYPE_OTHER, 2}, {"U16", FB2_DTYPE_OTHER, 2}, {"I8", FB2_DTYPE_OTHER, 1}, {"U8", FB2_DTYPE_OTHER, 1},
      {"BOOL", FB2_DTYPE_OTHER, 1}, {"F8_E4M3", FB2_DTYPE_OTHER, 1}, {"F8_E5M2", FB2_DTYPE_OTHER, 1}};
  for (auto &t : kTable)
    if (s == t.name) { *elem_bytes = t.bytes; return t.code; }
  *elem_bytes = 0;
  return -100;
}

struct TensorMeta {          // native_safetensors.rs:52-57
  int shard = 0;
  int dtype = FB2_DTYPE_OTHER;
  std::string dtype_name;
  std::vector<size_t> shape;
  size_t begin = 0, end = 0;  // byte range inside the shard's mapping
  size_t elems() const {
    size_t n = 1;
    for (size_t d : shape) n *= d;
    return n;
  }
};

// One shard file: mapping + the tensors its header lists (Shard::open, :69-116).
int open_shard(const std::string &path, int shard_idx, Mapping &map,
               std::vector<std::pair<std::string, TensorMeta>> &tensors) {
  const int fd = open(path.c_str(), O_RDONLY);
  if (fd < 0) { set_error("open \"%s\": %s", path.c_str(), strerror(errno)); return FB2_ERR_IO; }
  struct stat st;
  if (fstat(fd, &st) != 0) { set_error("stat \"%s\": %s", path.c_str(), strerror(errno)); close(fd); return FB2_ERR_IO; }
  const size_t len = (size_t)st.st_size;
  if (len < 8) { set_error("parse \"%s\": file is smaller than the 8-byte header length", path.c_str()); close(fd); return FB2_ERR_MODEL; }
  void *m = mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (m == MAP_FAILED) { set_error("mmap \"%s\": %s", path.c_str(), strerror(errno)); return FB2_ERR_IO; }
  map.base = (const uint8_t *)m;
  map.len = len;
  uint64_t hlen = 0;
  for (int i = 7; i >= 0; i--) hlen = (hlen << 8) | map.base[i];        // little-endian
  if (hlen > len - 8 || hlen > (100u << 20)) {
    set_error("parse \"%s\": header length %llu exceeds the file", path.c_str(), (unsigned long long)hlen);
    return FB2_ERR_MODEL;
  }
  const size_t data_base = 8 + (size_t)hlen, data_len = len - data_base;
  Json root;
  std::string err;
  if (!parse_json((const char *)map.base + 8, (size_t)hlen, root, err) || root.kind != Json::Obj) {
    set_error("parse \"%s\": header is not a JSON object (%s)", path.c_str(), err.c_str());
    return FB2_ERR_MODEL;
  }
  for (auto &kv : root.obj) {
    if (kv.first == "__metadata__") continue;
    const Json &t = kv.second;
    const Json *jd = t.get("dtype"), *js = t.get("shape"), *jo = t.get("data_offsets");
    if (!jd || jd->kind != Json::Str || !js || js->kind != Json::Arr || !jo || jo->kind != Json::Arr ||
        jo->arr.size() != 2 || !jo->arr[0].is_int || !jo->arr[1].is_int) {
      set_error("parse \"%s\": tensor '%s' has a malformed header entry", path.c_str(), kv.first.c_str());
      return FB2_ERR_MODEL;
    }
    TensorMeta tm;
    tm.shard = shard_idx;
    tm.dtype_name = jd->str;
    size_t eb = 0;
    tm.dtype = dtype_code(jd->str, &eb);
    if (tm.dtype == -100) {
      set_error("parse \"%s\": tensor '%s' has unknown dtype '%s'", path.c_str(), kv.first.c_str(), jd->str.c_str());
      return FB2_ERR_MODEL;
    }
    size_t n = 1;
    for (auto &d : js->arr) {
      if (!d.is_int || d.inum < 0) { set_error("parse \"%s\": tensor '%s' has a bad shape", path.c_str(), kv.first.c_str()); return FB2_ERR_MODEL; }
      tm.shape.push_back((size_t)d.inum);
      if (d.inum && n > SIZE_MAX / (size_t)d.inum) { set_error("parse \"%s\": tensor '%s' shape overflows", path.c_str(), kv.first.c_str()); return FB2_ERR_MODEL; }
      n *= (size_t)d.inum;
    }
    const long long b = jo->arr[0].inum, e = jo->arr[1].inum;
    if (b < 0 || e < b || (size_t)e > data_len || (size_t)(e - b) != n * eb) {
      set_error("parse \"%s\": tensor '%s' data_offsets [%lld, %lld) do not match %zu x %zu bytes inside a %zu-byte blob",
                path.c_str(), kv.first.c_str(), b, e, n, eb, data_len);
      return FB2_ERR_MODEL;
    }
    tm.begin = data_base + (size_t)b;
    tm.end = data_base + (size_t)e;
    tensors.emplace_back(kv.first, std::move(tm));
  }
  return FB2_OK;
}

// ------------------------------------------------------------------------ quant config
std::string lower(std::string s) {
  for (auto &c : s) c = (char)tolower((unsigned char)c);
  return s;
}
int method_code(const std::string &m) {
  const std::string l = lower(m);
  if (l == "gptq") return FB2_QUANT_GPTQ;
  if (l == "awq") return FB2_QUANT_AWQ;
  if (l == "gguf") return FB2_QUANT_GGUF;
  if (l == "none") return FB2_QUANT_NONE;
  return -1;
}

// load_quantize_config (:1475-1530): quantize_config.json first (AutoGPTQ), else the
// "quantization_config" block of config.json (transformers style).
int load_quant_config(const std::string &dir, fb2_quant_config_t *qc) {
  *qc = fb2_quant_config_t{};
  std::string text, err;
  const std::string p = dir + "/quantize_config.json";
  if (file_exists(p)) {
    int rc = read_text_file(p, text);
    if (rc) return rc;
    Json root;
    if (!parse_json(text.data(), text.size(), root, err) || root.kind != Json::Obj) {
      set_error("parse quantize_config.json: %s", err.empty() ? "not an object" : err.c_str());
      return FB2_ERR_MODEL;
    }
    // serde: `quant_method` with alias `method`, default None; an unknown variant is an error
    // (config.rs:10-22, rename_all = "lowercase": the match is case-sensitive)
    const Json *jm = root.get("quant_method");
    if (!jm) jm = root.get("method");
    int method = FB2_QUANT_NONE;
    if (jm) {
      method = jm->kind == Json::Str && jm->str == lower(jm->str) ? method_code(jm->str) : -1;
      if (method < 0) { set_error("parse quantize_config.json: unknown quant_method"); return FB2_ERR_MODEL; }
    }
    auto uint_field = [&](const char *key, unsigned long long *out) {
      const Json *j = root.get(key);
      if (!j) return true;
      if (j->kind != Json::Num || !j->is_int || j->inum < 0) return false;   // serde u32/usize reject negatives
      *out = (unsigned long long)j->inum;
      return true;
    };
    auto bool_field = [&](const char *key, int *out) {
      const Json *j = root.get(key);
      if (!j) return true;
      if (j->kind != Json::Bool) return false;
      *out = j->b ? 1 : 0;
      return true;
    };
    unsigned long long bits = 0, gs = 0;
    int desc = 0, sym = 0;
    if (!uint_field("bits", &bits) || !uint_field("group_size", &gs) || !bool_field("desc_act", &desc) ||
        !bool_field("sym", &sym)) {
      set_error("parse quantize_config.json: bits/group_size must be non-negative integers, desc_act/sym booleans");
      return FB2_ERR_MODEL;
    }
    qc->present = 1; qc->method = method; qc->bits = (uint32_t)bits; qc->group_size = (size_t)gs;
    qc->desc_act = desc; qc->sym = sym;
    return FB2_OK;
  }
  const std::string c = dir + "/config.json";
  if (file_exists(c)) {
    int rc = read_text_file(c, text);
    if (rc) return rc;
    Json root;
    if (!parse_json(text.data(), text.size(), root, err)) { set_error("parse config.json: %s", err.c_str()); return FB2_ERR_MODEL; }
    const Json *q = root.get("quantization_config");
    if (q) {
      const Json *jm = q->get("quant_method");
      int method = jm && jm->kind == Json::Str ? method_code(jm->str) : FB2_QUANT_NONE;   // lower-cased match (:1499)
      if (method < 0) method = FB2_QUANT_NONE;
      if (method != FB2_QUANT_NONE) {
        const Json *jb = q->get("bits"), *jg = q->get("group_size"), *jd = q->get("desc_act"), *js = q->get("sym");
        qc->present = 1;
        qc->method = method;
        qc->bits = jb && jb->is_int && jb->inum >= 0 ? (uint32_t)jb->inum : 0;
        long long g = jg && jg->is_int ? jg->inum : 128;
        qc->group_size = (size_t)std::max(0LL, g);
        qc->desc_act = jd && jd->kind == Json::Bool && jd->b;
        qc->sym = js && js->kind == Json::Bool && js->b;
      }
    }
  }
  return FB2_OK;
}

}  // namespace
}  // namespace fb2

using namespace fb2;

// ------------------------------------------------------------------------- the handle
struct fb2_ckpt {
  std::string dir;
  std::vector<std::unique_ptr<Mapping>> shards;
  std::vector<std::string> names;                         // shard order, header order
  std::unordered_map<std::string, TensorMeta> index;      // later shards win (HashMap insert, :161-163)
  fb2_quant_config_t qcfg{};

  const TensorMeta *find(const std::string &name) const {
    auto it = index.find(name);
    return it == index.end() ? nullptr : &it->second;
  }
  bool has(const std::string &name) const { return index.count(name) != 0; }
  const uint8_t *data(const TensorMeta &m) const { return shards[m.shard]->base + m.begin; }
};

namespace fb2 {
namespace {

int need(const fb2_ckpt *c, const std::string &name, const TensorMeta **out) {
  *out = c->find(name);
  if (!*out) { set_error("tensor '%s' not in index", name.c_str()); return FB2_ERR_MODEL; }
  return FB2_OK;
}

// read_f32 / dtype_to_f32 (:197-205, :1422-1473): F32 copied, F16 / BF16 widened.
int read_f32(const fb2_ckpt *c, const std::string &name, std::vector<float> &out, std::vector<size_t> &shape) {
  const TensorMeta *m;
  int rc = need(c, name, &m);
  if (rc) return rc;
  const size_t n = m->elems();
  out.resize(n);
  shape = m->shape;
  const uint8_t *src = c->data(*m);
  if (m->dtype == FB2_DTYPE_F32) {
    memcpy(out.data(), src, n * 4);
  } else if (m->dtype == FB2_DTYPE_F16) {
    for (size_t i = 0; i < n; i++) {
      uint16_t h;
      memcpy(&h, src + 2 * i, 2);
      out[i] = __half2float(__ushort_as_half(h));
    }
  } else if (m->dtype == FB2_DTYPE_BF16) {
    for (size_t i = 0; i < n; i++) {
      uint16_t h;
      memcpy(&h, src + 2 * i, 2);
      const uint32_t u = (uint32_t)h << 16;
      memcpy(&out[i], &u, 4);
    }
  } else {
    set_error("'%s': dtype %s not supported by the f32 read path", name.c_str(), m->dtype_name.c_str());
    return FB2_ERR_MODEL;
  }
  return FB2_OK;
}

// read_i32 (:305-329): qweight / qzeros / g_idx must be stored as I32.
int read_i32(const fb2_ckpt *c, const std::string &name, std::vector<int32_t> &out, std::vector<size_t> &shape) {
  const TensorMeta *m;
  int rc = need(c, name, &m);
  if (rc) return rc;
  if (m->dtype != FB2_DTYPE_I32) {
    set_error("'%s': expected I32, got %s", name.c_str(), m->dtype_name.c_str());
    return FB2_ERR_MODEL;
  }
  out.resize(m->elems());
  shape = m->shape;
  memcpy(out.data(), c->data(*m), out.size() * 4);
  return FB2_OK;
}

int require_gptq_config(const fb2_ckpt *c, const char *what) {
  if (!c->qcfg.present) {
    set_error("%s requires quantize_config.json (or config.json quantization_config)", what);
    return FB2_ERR_MODEL;
  }
  if (c->qcfg.method != FB2_QUANT_GPTQ) {
    set_error("%s: quant_method is not gptq", what);
    return FB2_ERR_MODEL;
  }
  return FB2_OK;
}

// canonicalize_gptq_qzeros_for_sym (:1242-1246)
void canonicalize_sym(const fb2_quant_config_t &q, std::vector<int32_t> &qzeros) {
  if (q.method == FB2_QUANT_GPTQ && q.sym && q.bits == 4) std::fill(qzeros.begin(), qzeros.end(), 0x77777777);
}

// validate_gptq_g_idx (:1288-1325): returns desc_act (non-trivial order) through *desc.
int validate_g_idx(const std::string &name, const fb2_quant_config_t &q, const std::vector<int32_t> *g_idx, size_t k,
                   int *desc) {
  *desc = 0;
  if (q.desc_act && !g_idx) {
    set_error("%s: quantize_config desc_act=true but no g_idx tensor was found", name.c_str());
    return FB2_ERR_MODEL;
  }
  if (!g_idx) return FB2_OK;
  if (q.group_size == 0) { set_error("%s: GPTQ g_idx present but group_size is 0", name.c_str()); return FB2_ERR_MODEL; }
  if (g_idx->size() != k) {
    set_error("%s: g_idx length %zu must match K=%zu", name.c_str(), g_idx->size(), k);
    return FB2_ERR_MODEL;
  }
  const size_t groups = (k + q.group_size - 1) / q.group_size;
  for (size_t i = 0; i < k; i++) {
    const int32_t g = (*g_idx)[i];
    if (g < 0 || (size_t)g >= groups) {
      set_error("%s: g_idx[%zu]=%d outside expected group range 0..%zu", name.c_str(), i, g, groups ? groups - 1 : 0);
      return FB2_ERR_MODEL;
    }
    if ((size_t)g != i / q.group_size) *desc = 1;      // gptq_g_idx_is_desc_act (:1134-1139)
  }
  return FB2_OK;
}

// One GPTQ linear as load_gptq receives it: the single-tensor case (:763-880) or the fusion of
// several parts along N (:887-1102).
struct GptqHost {
  std::vector<int32_t> qweight, qzeros, g_idx;
  std::vector<float> scales, bias;
  std::vector<size_t> part_n;          // widths of the fused parts (TP column split segments)
  size_t k = 0, n = 0;
  bool has_g_idx = false, has_bias = false;
  int desc_act = 0;
};

bool ends_with(const std::string &s, const char *suf, std::string *prefix) {
  const size_t l = strlen(suf);
  if (s.size() < l || s.compare(s.size() - l, l, suf) != 0) return false;
  *prefix = s.substr(0, s.size() - l);
  return true;
}

// load_linear's GPTQ resolution (:597-634): `<name>.qweight`, else the fusion shims.
int resolve_parts(const fb2_ckpt *c, const std::string &name, std::vector<std::string> &parts) {
  parts.clear();
  if (c->has(name + ".qweight")) { parts.push_back(name); return FB2_OK; }
  static const struct { const char *fused; const char *parts[3]; } kShims[] = {
      {"qkv_proj", {"q_proj", "k_proj", "v_proj"}},
      {"gate_up_proj", {"gate_proj", "up_proj", nullptr}},
      {"in_proj_qkvz", {"in_proj_qkv", "in_proj_z", nullptr}},
      {"in_proj_ba", {"in_proj_b", "in_proj_a", nullptr}}};
  for (auto &s : kShims) {
    std::string prefix;
    if (!ends_with(name, s.fused, &prefix)) continue;
    std::vector<std::string> cand;
    for (const char *p : s.parts)
      if (p) cand.push_back(prefix + p);
    bool all = true;
    for (auto &p : cand) all = all && c->has(p + ".qweight");
    if (all) { parts = cand; return FB2_OK; }
  }
  set_error("linear '%s': no '%s.qweight' and no fusable GPTQ parts in the checkpoint", name.c_str(), name.c_str());
  return FB2_ERR_MODEL;
}

int read_gptq(const fb2_ckpt *c, const std::string &name, bool want_data, GptqHost &h) {
  int rc = require_gptq_config(c, ("'" + name + "'").c_str());
  if (rc) return rc;
  const fb2_quant_config_t &q = c->qcfg;
  std::vector<std::string> parts;
  if ((rc = resolve_parts(c, name, parts))) return rc;
  const bool fused = parts.size() > 1;

  struct Part { std::vector<int32_t> qw, qz; std::vector<float> sc; size_t n, n_sc, n_qz; };
  std::vector<Part> ps(parts.size());
  size_t qw_rows = 0, sc_rows = 0, qz_rows = 0;
  std::vector<bool> has_gidx(parts.size(), false);
  for (size_t i = 0; i < parts.size(); i++) {
    const std::string &p = parts[i];
    std::vector<size_t> qs, ss, zs;
    if ((rc = read_i32(c, p + ".qweight", ps[i].qw, qs))) return rc;
    if ((rc = read_f32(c, p + ".scales", ps[i].sc, ss))) return rc;
    if ((rc = read_i32(c, p + ".qzeros", ps[i].qz, zs))) return rc;
    canonicalize_sym(q, ps[i].qz);
    if (qs.size() != 2 || ss.size() != 2 || zs.size() != 2) {
      set_error("GPTQ '%s': expected 2D qweight / scales / qzeros", p.c_str());
      return FB2_ERR_MODEL;
    }
    if (i == 0) { qw_rows = qs[0]; sc_rows = ss[0]; qz_rows = zs[0]; }
    else if (qs[0] != qw_rows || ss[0] != sc_rows || zs[0] != qz_rows) {
      set_error("GPTQ fusion row mismatch on '%s'", p.c_str());
      return FB2_ERR_MODEL;
    }
    ps[i].n = qs[1]; ps[i].n_sc = ss[1]; ps[i].n_qz = zs[1];
    // what the device ingest indexes: scales [K/g, n], qzeros [K/g, n/8] (the reference checks the
    // scales width, :838-842; the rest would be an out-of-bounds read in load_gptq)
    if (ss[1] != qs[1] || zs[1] * 8 != qs[1] || ss[0] != zs[0]) {
      set_error("'%s.scales' [%zu, %zu] / qzeros [%zu, %zu] incompatible with qweight [%zu, %zu]", p.c_str(), ss[0], ss[1],
                zs[0], zs[1], qs[0], qs[1]);
      return FB2_ERR_MODEL;
    }
    const std::string gk = p + ".g_idx";
    if (c->has(gk)) {
      std::vector<int32_t> gx;
      std::vector<size_t> gs;
      if ((rc = read_i32(c, gk, gx, gs))) return rc;
      if (fused && (gs.size() != 1 || gs[0] != qw_rows * 8)) {
        set_error("GPTQ fusion '%s': g_idx shape incompatible with K=%zu", p.c_str(), qw_rows * 8);
        return FB2_ERR_MODEL;
      }
      if (!h.has_g_idx) { h.g_idx = std::move(gx); h.has_g_idx = true; }
      else if (h.g_idx != gx) {
        set_error("GPTQ fusion '%s': g_idx mismatch with first part; fused qkv/gate_up requires identical act-order across parts",
                  p.c_str());
        return FB2_ERR_MODEL;
      }
      has_gidx[i] = true;
    }
    h.part_n.push_back(ps[i].n);
    h.n += ps[i].n;
  }
  h.k = qw_rows * 8;
  if (h.has_g_idx)
    for (size_t i = 0; i < parts.size(); i++)
      if (!has_gidx[i]) {
        set_error("GPTQ fusion requires all parts to carry g_idx when any part does; missing g_idx for '%s'", parts[i].c_str());
        return FB2_ERR_MODEL;
      }
  std::string label = fused ? "GPTQ fusion " : "";
  for (size_t i = 0; i < parts.size(); i++) label += (i ? "+" : "") + parts[i];
  if ((rc = validate_g_idx(label, q, h.has_g_idx ? &h.g_idx : nullptr, h.k, &h.desc_act))) return rc;
  if (q.group_size == 0 || h.k % q.group_size || sc_rows != h.k / q.group_size) {
    set_error("%s: %zu scale rows do not match K=%zu / group_size=%zu", label.c_str(), sc_rows, h.k, q.group_size);
    return FB2_ERR_MODEL;
  }

  // biases: all parts or none (:1058-1084); the single linear checks the length (:850-857)
  size_t with_bias = 0;
  for (auto &p : parts) with_bias += c->has(p + ".bias") ? 1 : 0;
  if (with_bias && with_bias != parts.size()) {
    set_error("GPTQ fusion: inconsistent bias presence across parts");
    return FB2_ERR_MODEL;
  }
  h.has_bias = with_bias != 0;
  if (h.has_bias) {
    for (auto &p : parts) {
      std::vector<float> b;
      std::vector<size_t> bs;
      if ((rc = read_f32(c, p + ".bias", b, bs))) return rc;
      h.bias.insert(h.bias.end(), b.begin(), b.end());
    }
    if (h.bias.size() != h.n) {
      set_error("%s: bias length %zu != out_features %zu", label.c_str(), h.bias.size(), h.n);
      return FB2_ERR_MODEL;
    }
  }
  if (!want_data) return FB2_OK;

  // row-major concatenation along N: for each row, every part's columns (:970-988)
  if (!fused) {
    h.qweight = std::move(ps[0].qw); h.scales = std::move(ps[0].sc); h.qzeros = std::move(ps[0].qz);
    return FB2_OK;
  }
  h.qweight.reserve(qw_rows * h.n);
  for (size_t r = 0; r < qw_rows; r++)
    for (auto &p : ps) h.qweight.insert(h.qweight.end(), p.qw.begin() + r * p.n, p.qw.begin() + (r + 1) * p.n);
  h.scales.reserve(sc_rows * h.n);
  for (size_t r = 0; r < sc_rows; r++)
    for (auto &p : ps) h.scales.insert(h.scales.end(), p.sc.begin() + r * p.n_sc, p.sc.begin() + (r + 1) * p.n_sc);
  h.qzeros.reserve(qz_rows * h.n / 8);
  for (size_t r = 0; r < qz_rows; r++)
    for (auto &p : ps) h.qzeros.insert(h.qzeros.end(), p.qz.begin() + r * p.n_qz, p.qz.begin() + (r + 1) * p.n_qz);
  return FB2_OK;
}

// load_stacked_gptq_experts (:388-585): experts [e0, e0 + count) of one layer, each the row-major
// concatenation of `projs` along N; one shared g_idx.
struct StackHost {
  std::vector<std::vector<int32_t>> qweight, qzeros;
  std::vector<std::vector<float>> scales;
  std::vector<int32_t> g_idx;
  bool has_g_idx = false;
  size_t k = 0, n_per_expert = 0;
  int desc_act = 0;
};

std::string expert_prefix(const std::string &fmt, size_t e) {
  std::string out = fmt;
  const std::string num = std::to_string(e);
  size_t pos = 0;
  while ((pos = out.find("{e}", pos)) != std::string::npos) {     // str::replace: every occurrence
    out.replace(pos, 3, num);
    pos += num.size();
  }
  return out;
}

int read_stack(const fb2_ckpt *c, const std::string &fmt, size_t e0, size_t count, const std::vector<std::string> &projs,
               StackHost &s) {
  int rc = require_gptq_config(c, "load_stacked_gptq_experts");
  if (rc) return rc;
  if (count == 0 || projs.empty()) { set_error("load_stacked_gptq_experts: empty expert or projection list"); return FB2_ERR_MODEL; }
  const fb2_quant_config_t &q = c->qcfg;
  size_t qw_rows = 0, sc_rows = 0, qz_rows = 0, n_sc0 = 0, n_qz0 = 0;
  s.qweight.resize(count); s.scales.resize(count); s.qzeros.resize(count);
  for (size_t i = 0; i < count; i++) {
    const std::string prefix = expert_prefix(fmt, e0 + i);
    struct Part { std::vector<int32_t> qw, qz; std::vector<float> sc; size_t n, n_sc, n_qz; };
    std::vector<Part> ps(projs.size());
    size_t e_n = 0, e_sc = 0, e_qz = 0;
    for (size_t j = 0; j < projs.size(); j++) {
      const std::string name = prefix + projs[j];
      std::vector<size_t> qs, ss, zs;
      if ((rc = read_i32(c, name + ".qweight", ps[j].qw, qs))) return rc;
      if ((rc = read_f32(c, name + ".scales", ps[j].sc, ss))) return rc;
      if ((rc = read_i32(c, name + ".qzeros", ps[j].qz, zs))) return rc;
      canonicalize_sym(q, ps[j].qz);
      if (qs.size() != 2 || ss.size() != 2 || zs.size() != 2) {
        set_error("stacked GPTQ '%s': expected 2D tensors", name.c_str());
        return FB2_ERR_MODEL;
      }
      if (qw_rows == 0) { qw_rows = qs[0]; sc_rows = ss[0]; qz_rows = zs[0]; s.k = qs[0] * 8; }
      else if (qs[0] != qw_rows || ss[0] != sc_rows || zs[0] != qz_rows) {
        set_error("stacked GPTQ '%s': row mismatch qw %zu sc %zu qz %zu vs ref %zu/%zu/%zu", name.c_str(), qs[0], ss[0], zs[0],
                  qw_rows, sc_rows, qz_rows);
        return FB2_ERR_MODEL;
      }
      if (ss[1] != qs[1] || zs[1] * 8 != qs[1]) {
        set_error("stacked GPTQ '%s': scales / qzeros widths %zu / %zu do not match qweight width %zu", name.c_str(), ss[1], zs[1], qs[1]);
        return FB2_ERR_MODEL;
      }
      ps[j].n = qs[1]; ps[j].n_sc = ss[1]; ps[j].n_qz = zs[1];
      e_n += qs[1]; e_sc += ss[1]; e_qz += zs[1];
      const std::string gk = name + ".g_idx";
      if (c->has(gk)) {
        std::vector<int32_t> gx;
        std::vector<size_t> gs;
        if ((rc = read_i32(c, gk, gx, gs))) return rc;
        if (!s.has_g_idx) { s.g_idx = std::move(gx); s.has_g_idx = true; }
        else if (s.g_idx != gx) {
          set_error("stacked GPTQ '%s': g_idx mismatch with first expert -- one act-order is required across the experts of a stack",
                    name.c_str());
          return FB2_ERR_MODEL;
        }
      }
    }
    if (i == 0) { s.n_per_expert = e_n; n_sc0 = e_sc; n_qz0 = e_qz; }
    else if (e_n != s.n_per_expert || e_sc != n_sc0 || e_qz != n_qz0) {
      set_error("stacked GPTQ expert %zu N mismatch: qw %zu sc %zu qz %zu vs expert %zu %zu/%zu/%zu", e0 + i, e_n, e_sc, e_qz, e0,
                s.n_per_expert, n_sc0, n_qz0);
      return FB2_ERR_MODEL;
    }
    auto &qw = s.qweight[i]; auto &sc = s.scales[i]; auto &qz = s.qzeros[i];
    qw.reserve(qw_rows * e_n); sc.reserve(sc_rows * e_sc); qz.reserve(qz_rows * e_qz);
    for (size_t r = 0; r < qw_rows; r++)
      for (auto &p : ps) qw.insert(qw.end(), p.qw.begin() + r * p.n, p.qw.begin() + (r + 1) * p.n);
    for (size_t r = 0; r < sc_rows; r++)
      for (auto &p : ps) sc.insert(sc.end(), p.sc.begin() + r * p.n_sc, p.sc.begin() + (r + 1) * p.n_sc);
    for (size_t r = 0; r < qz_rows; r++)
      for (auto &p : ps) qz.insert(qz.end(), p.qz.begin() + r * p.n_qz, p.qz.begin() + (r + 1) * p.n_qz);
  }
  if ((rc = validate_g_idx("stacked GPTQ experts", q, s.has_g_idx ? &s.g_idx : nullptr, s.k, &s.desc_act))) return rc;
  if (q.group_size == 0 || s.k % q.group_size || sc_rows != s.k / q.group_size || qz_rows != sc_rows) {
    set_error("stacked GPTQ experts: %zu scale rows do not match K=%zu / group_size=%zu", sc_rows, s.k, q.group_size);
    return FB2_ERR_MODEL;
  }
  return FB2_OK;
}

template <class T>
void copy_out(T *dst, const std::vector<T> &src) {
  if (dst && !src.empty()) memcpy(dst, src.data(), src.size() * sizeof(T));
}

}  // namespace
}  // namespace fb2

// ------------------------------------------------------------------------------ C ABI
extern "C" {

int fb2_ckpt_open(const char *model_dir, fb2_ckpt_t **out) {
  if (!out) { set_error("out is NULL"); return FB2_ERR_ARG; }
  *out = nullptr;
  if (!model_dir) { set_error("model_dir is NULL"); return FB2_ERR_ARG; }
  std::unique_ptr<fb2_ckpt> c(new fb2_ckpt());
  c->dir = model_dir;
  while (c->dir.size() > 1 && c->dir.back() == '/') c->dir.pop_back();
  std::vector<std::string> files;
  if (file_exists(c->dir + "/model.safetensors")) {
    files.push_back("model.safetensors");
  } else if (file_exists(c->dir + "/model.safetensors.index.json")) {
    // parse_sharded_index (:178-194): the distinct file names of weight_map, sorted
    std::string text, err;
    int rc = read_text_file(c->dir + "/model.safetensors.index.json", text);
    if (rc) return rc;
    Json root;
    if (!parse_json(text.data(), text.size(), root, err)) { set_error("index json: %s", err.c_str()); return FB2_ERR_MODEL; }
    const Json *wm = root.get("weight_map");
    if (!wm || wm->kind != Json::Obj) { set_error("index missing weight_map"); return FB2_ERR_MODEL; }
    for (auto &kv : wm->obj)
      if (kv.second.kind == Json::Str) files.push_back(kv.second.str);
    std::sort(files.begin(), files.end());
    files.erase(std::unique(files.begin(), files.end()), files.end());
  } else {
    set_error("no safetensors files in \"%s\"", c->dir.c_str());
    return FB2_ERR_MODEL;
  }
  for (size_t i = 0; i < files.size(); i++) {
    std::unique_ptr<Mapping> map(new Mapping());
    std::vector<std::pair<std::string, TensorMeta>> tensors;
    int rc = open_shard(c->dir + "/" + files[i], (int)i, *map, tensors);
    if (rc) return rc;
    c->shards.push_back(std::move(map));
    for (auto &t : tensors) {
      if (!c->index.count(t.first)) c->names.push_back(t.first);
      c->index[t.first] = std::move(t.second);
    }
  }
  int rc = load_quant_config(c->dir, &c->qcfg);
  if (rc) return rc;
  *out = c.release();
  return FB2_OK;
}

void fb2_ckpt_close(fb2_ckpt_t *ckpt) { delete ckpt; }

int fb2_ckpt_quant_config(const fb2_ckpt_t *ckpt, fb2_quant_config_t *out) {
  if (!ckpt || !out) { set_error("NULL argument"); return FB2_ERR_ARG; }
  *out = ckpt->qcfg;
  return FB2_OK;
}

size_t fb2_ckpt_num_tensors(const fb2_ckpt_t *ckpt) { return ckpt ? ckpt->names.size() : 0; }
const char *fb2_ckpt_tensor_name(const fb2_ckpt_t *ckpt, size_t i) {
  return ckpt && i < ckpt->names.size() ? ckpt->names[i].c_str() : nullptr;
}
int fb2_ckpt_has_tensor(const fb2_ckpt_t *ckpt, const char *name) { return ckpt && name && ckpt->has(name) ? 1 : 0; }

int fb2_ckpt_tensor_info(const fb2_ckpt_t *ckpt, const char *name, int *dtype, size_t *rank, size_t *shape,
                         size_t *nbytes) {
  if (!ckpt || !name) { set_error("NULL argument"); return FB2_ERR_ARG; }
  const TensorMeta *m;
  int rc = need(ckpt, name, &m);
  if (rc) return rc;
  if (m->shape.size() > FB2_MAX_RANK) { set_error("'%s': rank %zu exceeds %d", name, m->shape.size(), FB2_MAX_RANK); return FB2_ERR_UNSUPPORTED; }
  if (dtype) *dtype = m->dtype;
  if (rank) *rank = m->shape.size();
  if (shape)
    for (size_t i = 0; i < m->shape.size(); i++) shape[i] = m->shape[i];
  if (nbytes) *nbytes = m->end - m->begin;
  return FB2_OK;
}

int fb2_ckpt_read_f32(const fb2_ckpt_t *ckpt, const char *name, float *out, size_t capacity) {
  if (!ckpt || !name || !out) { set_error("NULL argument"); return FB2_ERR_ARG; }
  std::vector<float> v;
  std::vector<size_t> shape;
  int rc = read_f32(ckpt, name, v, shape);
  if (rc) return rc;
  if (v.size() > capacity) { set_error("'%s': %zu elements do not fit a buffer of %zu", name, v.size(), capacity); return FB2_ERR_SHAPE; }
  copy_out(out, v);
  return FB2_OK;
}

int fb2_ckpt_read_i32(const fb2_ckpt_t *ckpt, const char *name, int32_t *out, size_t capacity) {
  if (!ckpt || !name || !out) { set_error("NULL argument"); return FB2_ERR_ARG; }
  std::vector<int32_t> v;
  std::vector<size_t> shape;
  int rc = read_i32(ckpt, name, v, shape);
  if (rc) return rc;
  if (v.size() > capacity) { set_error("'%s': %zu elements do not fit a buffer of %zu", name, v.size(), capacity); return FB2_ERR_SHAPE; }
  copy_out(out, v);
  return FB2_OK;
}

int fb2_ckpt_gptq_info(const fb2_ckpt_t *ckpt, const char *name, fb2_gptq_info_t *info) {
  if (!ckpt || !name || !info) { set_error("NULL argument"); return FB2_ERR_ARG; }
  GptqHost h;
  int rc = read_gptq(ckpt, name, false, h);
  if (rc) return rc;
  *info = fb2_gptq_info_t{};
  info->k = h.k; info->n = h.n;
  info->has_g_idx = h.has_g_idx; info->has_bias = h.has_bias; info->desc_act = h.desc_act;
  info->n_parts = (int)h.part_n.size();
  for (size_t i = 0; i < h.part_n.size() && i < 4; i++) info->part_n[i] = h.part_n[i];
  return FB2_OK;
}

int fb2_ckpt_read_gptq(const fb2_ckpt_t *ckpt, const char *name, int32_t *qweight, float *scales, int32_t *qzeros,
                       int32_t *g_idx, float *bias) {
  if (!ckpt || !name) { set_error("NULL argument"); return FB2_ERR_ARG; }
  GptqHost h;
  int rc = read_gptq(ckpt, name, true, h);
  if (rc) return rc;
  copy_out(qweight, h.qweight);
  copy_out(scales, h.scales);
  copy_out(qzeros, h.qzeros);
  if (h.has_g_idx) copy_out(g_idx, h.g_idx);
  if (h.has_bias) copy_out(bias, h.bias);
  return FB2_OK;
}

int fb2_ckpt_load_gptq_linear(const fb2_ckpt_t *ckpt, const char *name, int device, fb2_linear_t **out) {
  if (!out) { set_error("out is NULL"); return FB2_ERR_ARG; }
  *out = nullptr;
  if (!ckpt || !name) { set_error("NULL argument"); return FB2_ERR_ARG; }
  GptqHost h;
  int rc = read_gptq(ckpt, name, true, h);
  if (rc) return rc;
  return fb2_gptq_linear_create(h.qweight.data(), h.scales.data(), h.qzeros.data(), h.has_g_idx ? h.g_idx.data() : nullptr,
                                h.has_bias ? h.bias.data() : nullptr, ckpt->qcfg.bits, ckpt->qcfg.group_size, h.k, h.n,
                                device, out);
}

int fb2_ckpt_load_gptq_linear_sharded(const fb2_ckpt_t *ckpt, const char *name, int device, int tp_rank, int tp_size,
                                      int split, int32_t *x_cols_out, fb2_linear_t **out) {
  if (!out) { set_error("out is NULL"); return FB2_ERR_ARG; }
  *out = nullptr;
  if (!ckpt || !name) { set_error("NULL argument"); return FB2_ERR_ARG; }
  GptqHost h;
  int rc = read_gptq(ckpt, name, true, h);
  if (rc) return rc;
  // column split: the fused parts (q|k|v, gate|up) are the segments that are split separately
  return fb2_gptq_linear_create_sharded(h.qweight.data(), h.scales.data(), h.qzeros.data(),
                                        h.has_g_idx ? h.g_idx.data() : nullptr, h.has_bias ? h.bias.data() : nullptr,
                                        ckpt->qcfg.bits, ckpt->qcfg.group_size, h.k, h.n, device, tp_rank, tp_size, split,
                                        h.part_n.size() > 1 ? h.part_n.data() : nullptr,
                                        h.part_n.size() > 1 ? (int)h.part_n.size() : 0, x_cols_out, out);
}

int fb2_ckpt_stacked_gptq_info(const fb2_ckpt_t *ckpt, const char *expert_prefix_fmt, const char *const *proj_names,
                               size_t n_proj, size_t expert, size_t *k, size_t *n_per_expert, int *has_g_idx) {
  if (!ckpt || !expert_prefix_fmt || !proj_names) { set_error("NULL argument"); return FB2_ERR_ARG; }
  int rc = require_gptq_config(ckpt, "load_stacked_gptq_experts");
  if (rc) return rc;
  size_t kk = 0, n = 0;
  int g = 0;
  const std::string prefix = expert_prefix(expert_prefix_fmt, expert);
  for (size_t j = 0; j < n_proj; j++) {
    const TensorMeta *m;
    const std::string name = prefix + proj_names[j];
    if ((rc = need(ckpt, name + ".qweight", &m))) return rc;
    if (m->shape.size() != 2) { set_error("stacked GPTQ '%s': expected 2D qweight", name.c_str()); return FB2_ERR_MODEL; }
    kk = m->shape[0] * 8;
    n += m->shape[1];
    g |= ckpt->has(name + ".g_idx") ? 1 : 0;
  }
  if (k) *k = kk;
  if (n_per_expert) *n_per_expert = n;
  if (has_g_idx) *has_g_idx = g;
  return FB2_OK;
}

int fb2_ckpt_read_stacked_gptq_expert(const fb2_ckpt_t *ckpt, const char *expert_prefix_fmt, const char *const *proj_names,
                                      size_t n_proj, size_t expert, int32_t *qweight, float *scales, int32_t *qzeros,
                                      int32_t *g_idx) {
  if (!ckpt || !expert_prefix_fmt || !proj_names) { set_error("NULL argument"); return FB2_ERR_ARG; }
  std::vector<std::string> projs(proj_names, proj_names + n_proj);
  StackHost s;
  int rc = read_stack(ckpt, expert_prefix_fmt, expert, 1, projs, s);
  if (rc) return rc;
  copy_out(qweight, s.qweight[0]);
  copy_out(scales, s.scales[0]);
  copy_out(qzeros, s.qzeros[0]);
  if (s.has_g_idx) copy_out(g_idx, s.g_idx);
  return FB2_OK;
}

int fb2_ckpt_load_stacked_gptq_experts(const fb2_ckpt_t *ckpt, const char *expert_prefix_fmt, size_t expert_begin,
                                       size_t num_experts, const char *const *proj_names, size_t n_proj, int device,
                                       fb2_stack_t **out, size_t *n_per_expert, size_t *k) {
  if (!out) { set_error("out is NULL"); return FB2_ERR_ARG; }
  *out = nullptr;
  if (!ckpt || !expert_prefix_fmt || !proj_names) { set_error("NULL argument"); return FB2_ERR_ARG; }
  std::vector<std::string> projs(proj_names, proj_names + n_proj);
  StackHost s;
  int rc = read_stack(ckpt, expert_prefix_fmt, expert_begin, num_experts, projs, s);
  if (rc) return rc;
  std::vector<const int32_t *> qw(num_experts), qz(num_experts);
  std::vector<const float *> sc(num_experts);
  for (size_t e = 0; e < num_experts; e++) { qw[e] = s.qweight[e].data(); sc[e] = s.scales[e].data(); qz[e] = s.qzeros[e].data(); }
  rc = fb2_gptq_stack_create(qw.data(), sc.data(), qz.data(), s.has_g_idx ? s.g_idx.data() : nullptr, ckpt->qcfg.bits,
                             ckpt->qcfg.group_size, s.k, s.n_per_expert, num_experts, device, out);
  if (rc) return rc;
  if (n_per_expert) *n_per_expert = s.n_per_expert;
  if (k) *k = s.k;
  return FB2_OK;
}

}  // extern "C"
