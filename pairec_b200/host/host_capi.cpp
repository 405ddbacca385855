// host_capi.cpp — C ABI over the host mirror so that tests (ctypes) can drive it the way /api/recommend drives the
// reference: one recconf JSON in, RecommendParam JSON in, RecommendResponse JSON out
// (web/recommend_controller.go:24-31,51-60,112-157).  "score" is added to each item for the parity checks.
#include <cstring>

#include "pairec_host.hpp"
#include <cmath>

using namespace pairec;

namespace {
thread_local std::string g_err;
struct UserVectors : recall::VectorDao {
  std::map<std::string, std::string> m;
  Error VectorString(const std::string& id, std::string* out) override {
    auto it = m.find(id);
    if (it == m.end()) return recall::VectoryEmptyError;
    *out = it->second;
    return "";
  }
};
}  // namespace

struct ph_server {
  recconf::RecommendConfig conf;
  std::shared_ptr<GpuCatalog> catalog = std::make_shared<GpuCatalog>();
  std::map<std::string, std::vector<module::ItemPtr>> context_items;  // recall name -> items
  std::shared_ptr<UserVectors> vectors = std::make_shared<UserVectors>();
  uint64_t next_id = 1;
};

extern "C" {

const char* ph_last_error(void) { return g_err.c_str(); }

int ph_create(const char* recconf_json, ph_server** out) {
  if (!recconf_json || !out) { g_err = "null argument"; return 1; }
  auto* s = new ph_server();
  Error e = recconf::LoadConfig(recconf_json, &s->conf);
  if (!e.empty()) { g_err = e; delete s; return 1; }
  ResetRegistries();
  algorithm::Load(s->conf);   // pairec.go:88-109 runBeforeStart order: algorithm.Load, then register(conf)
  recall::Load(s->conf);
  filter::Load(s->conf);
  sort::Load(s->conf);
  *out = s;
  return 0;
}
void ph_destroy(ph_server* s) { delete s; }

int ph_attach_engine(ph_server* s, void* prg) {
  if (!s) { g_err = "null server"; return 1; }
  s->catalog->h = static_cast<prg_handle*>(prg);
  return 0;
}
int ph_set_catalog_ids(ph_server* s, const char* const* ids, uint64_t n) {
  if (!s || (!ids && n)) { g_err = "null argument"; return 1; }
  std::vector<std::string> v;
  v.reserve(n);
  for (uint64_t i = 0; i < n; ++i) v.emplace_back(ids[i]);
  s->catalog->SetIds(std::move(v));
  return 0;
}
// items of an in-memory recall (config 1); properties_json = {"score": 0.3, "brand": "x", ...}
int ph_add_context_item(ph_server* s, const char* recall_name, const char* item_id, double score, const char* properties_json) {
  if (!s || !recall_name || !item_id) { g_err = "null argument"; return 1; }
  auto item = module::NewItem(item_id);
  item->Score = score;
  if (properties_json && *properties_json) {
    Json j;
    std::string err;
    if (!Json::parse(properties_json, &j, &err) || j.type != Json::Object) { g_err = "bad properties JSON: " + err; return 1; }
    for (auto& kv : j.obj) {
      if (kv.second.type == Json::Number) {
        if (kv.second.is_int) item->Properties[kv.first] = (int64_t)kv.second.num;
        else item->Properties[kv.first] = kv.second.num;
      } else if (kv.second.type == Json::String) item->Properties[kv.first] = kv.second.str;
    }
  }
  s->context_items[recall_name].push_back(item);
  return 0;
}
// makes the in-memory recalls and the sort strategies visible (call after the items are added)
int ph_commit(ph_server* s) {
  if (!s) { g_err = "null server"; return 1; }
  for (auto& kv : s->context_items)
    recall::RegisterRecall(kv.first, std::make_shared<recall::ContextItemRecall>(kv.first, kv.second));
  sort::Load(s->conf);
  return 0;
}
// GPU plugins, registered the way a user start hook would (pairec.go:58): names must match the recconf
int ph_register_gpu_plugins(ph_server* s, const char* recall_algo, const char* rank_algo, int model, const char* dpp_sort) {
  if (!s || !s->catalog->h) { g_err = "no engine attached"; return 1; }
  if (recall_algo && *recall_algo)
    algorithm::RegisterAlgorithm(recall_algo, std::make_shared<algorithm::GpuVectorAlgorithm>(s->catalog));
  if (rank_algo && *rank_algo) {
    algorithm::RegisterAlgorithm(rank_algo, std::make_shared<algorithm::GpuRankAlgorithm>(s->catalog, model));
    // the IAlgorithm route sees only feature maps: the scenes that name the algorithm get the LoadFeatureFunc that writes
    // the item id into the item's properties (feature.RegisterLoadFeatureFunc, service/feature/feature_service.go:36-38)
    for (auto& rc : s->conf.RankConf)
      for (auto& a : rc.second.RankAlgoList)
        if (a == rank_algo) feature::RegisterLoadFeatureFunc(rc.first, feature::ItemIdProperty());
  }
  if (dpp_sort && *dpp_sort) {
    recconf::DPPSortConfig dc;
    dc.Alpha = 1.0;
    for (auto& sc : s->conf.SortConfs)
      if (sc.Name == dpp_sort) dc = sc.DPPConf;
    sort::RegisterSort(dpp_sort, std::make_shared<sort::GpuDPPSort>(dc, s->catalog));
  }
  for (auto& sc : s->conf.SortConfs)  // SortType "SSDSort" entries become GPU-backed sorts under their own names
    if (sc.SortType == "SSDSort") sort::RegisterSort(sc.Name, std::make_shared<sort::GpuSSDSort>(sc.SSDConf, s->catalog));
  for (auto& rc : s->conf.RecallConfs)
    if (rc.RecallType == "VectorRecall") recall::RegisterRecall(rc.Name, std::make_shared<recall::VectorRecall>(rc, s->vectors));
  sort::Load(s->conf);
  return 0;
}
static std::vector<ingest::FieldSpec> parse_field_specs(const Json& spec) {
  std::vector<ingest::FieldSpec> specs;
  for (const Json& e : spec.arr) {
    ingest::FieldSpec fs;
    fs.Column = e["column"].as_string();
    fs.IsId = e["id"].as_bool(false);
    for (const Json& v : e["vocab"].arr) {
      if (v.type == Json::String) fs.Vocab.push_back(v.str);
      else if (v.type == Json::Number)
        fs.Vocab.push_back(ingest::ToString(v.num == std::floor(v.num) && std::fabs(v.num) < 9e15 ? module::Value((int64_t)v.num)
                                                                                                : module::Value(v.num)));
    }
    specs.push_back(std::move(fs));
  }
  return specs;
}
// The GPU rank as a rank.IRank of `scene` (rank.RegisterRank, service/rank/rank_service.go:45-60): one prg_rank_ex call
// per request with the Items and the User.  user_fields_json: field specs as in ph_encode_fields, applied to the
// request's user features; dense_columns_json: ["col", ...] numeric user features for the tower.
int ph_register_gpu_rank(ph_server* s, const char* scene, const char* name, int model, const char* user_fields_json,
                         const char* dense_columns_json, int heads) {
  if (!s || !s->catalog->h || !scene || !name) { g_err = "no engine attached / null argument"; return 1; }
  Json spec, dense;
  std::string err;
  if (!Json::parse(user_fields_json ? user_fields_json : "[]", &spec, &err) || spec.type != Json::Array) { g_err = "bad user field spec JSON: " + err; return 1; }
  if (!Json::parse(dense_columns_json ? dense_columns_json : "[]", &dense, &err) || dense.type != Json::Array) { g_err = "bad dense column JSON: " + err; return 1; }
  std::vector<std::string> cols;
  for (const Json& v : dense.arr) cols.push_back(v.as_string());
  rank::RegisterRank(scene, std::make_shared<rank::GpuRank>(s->catalog, name, model, parse_field_specs(spec), std::move(cols), heads));
  return 0;
}
// The GPU rank as an IAlgorithm for scenes whose RankConf.Processor is "EasyRec" (algorithm.RegisterAlgorithm): the stock
// RankService hands it one easyrec.PBRequest per batch — item ids + the request's user features (algo_data.go:292-325).
// outputs_json: the tower's head names in order (one name or none: single-score responses).
int ph_register_gpu_easyrec(ph_server* s, const char* algo_name, int model, const char* user_fields_json,
                            const char* dense_columns_json, const char* outputs_json) {
  if (!s || !s->catalog->h || !algo_name) { g_err = "no engine attached / null argument"; return 1; }
  Json spec, dense, outs;
  std::string err;
  if (!Json::parse(user_fields_json ? user_fields_json : "[]", &spec, &err) || spec.type != Json::Array) { g_err = "bad user field spec JSON: " + err; return 1; }
  if (!Json::parse(dense_columns_json ? dense_columns_json : "[]", &dense, &err) || dense.type != Json::Array) { g_err = "bad dense column JSON: " + err; return 1; }
  if (!Json::parse(outputs_json ? outputs_json : "[]", &outs, &err) || outs.type != Json::Array) { g_err = "bad outputs JSON: " + err; return 1; }
  std::vector<std::string> cols, names;
  for (const Json& v : dense.arr) cols.push_back(v.as_string());
  for (const Json& v : outs.arr) names.push_back(v.as_string());
  algorithm::RegisterAlgorithm(algo_name, std::make_shared<algorithm::GpuEasyrecAlgorithm>(s->catalog, model, parse_field_specs(spec),
                                                                                         std::move(cols), std::move(names)));
  return 0;
}

static void json_props(const Json& j, module::Features* out) {
  for (auto& kv : j.obj) {
    if (kv.second.type == Json::Number) {
      if (kv.second.is_int) (*out)[kv.first] = (int64_t)kv.second.num;
      else (*out)[kv.first] = kv.second.num;
    } else if (kv.second.type == Json::String) (*out)[kv.first] = kv.second.str;
  }
}
static std::string json_value(const module::Value& v) {
  if (auto d = std::get_if<double>(&v)) return Json::number(*d);
  if (auto i = std::get_if<int64_t>(&v)) return std::to_string(*i);
  return Json::quote(std::get<std::string>(v));
}
// rank::EasyrecAlgoDataGenerator as a pure function, for the CPU tests: RankConf.ContextFeatures / ItemFeatures, the items
// ([{"item_id": ..., "properties": {...}}, ...]) and the user's properties in, the PBRequests of the request's batches out
// (JSON list of {"user_features", "item_ids", "context_features", "item_features"}), as RankService would hand them to
// algorithm.Run.  Returns the bytes needed (incl. NUL); writes when it fits.
long long ph_easyrec_requests(const char* context_features_json, const char* item_features_json, const char* items_json,
                              const char* user_json, int batch_count, char* out, unsigned long long cap) {
  Json cf, itf, items, user;
  std::string err;
  if (!Json::parse(context_features_json ? context_features_json : "[]", &cf, &err) || cf.type != Json::Array ||
      !Json::parse(item_features_json ? item_features_json : "[]", &itf, &err) || itf.type != Json::Array ||
      !Json::parse(items_json ? items_json : "[]", &items, &err) || items.type != Json::Array ||
      !Json::parse(user_json ? user_json : "{}", &user, &err) || user.type != Json::Object) {
    g_err = "bad JSON argument: " + err;
    return -1;
  }
  std::vector<std::string> ctxNames, itemNames;
  for (const Json& v : cf.arr) ctxNames.push_back(v.as_string());
  for (const Json& v : itf.arr) itemNames.push_back(v.as_string());
  module::User u;
  json_props(user, &u.Properties);
  const module::Features userFeatures = u.MakeUserFeatures2();
  rank::EasyrecAlgoDataGenerator gen(ctxNames);
  gen.SetItemFeatures(itemNames);
  const bool want = !ctxNames.empty() || !itemNames.empty();
  if (batch_count <= 0) batch_count = 100;
  std::vector<rank::EasyrecAlgoDataGenerator::AlgoData> batches;
  int i = 0;
  for (const Json& e : items.arr) {
    auto it = module::NewItem(e["item_id"].as_string());
    json_props(e["properties"], &it->Properties);
    module::Features f;
    if (want) f = it->GetFeatures();
    gen.AddFeatures(it, want ? &f : nullptr, userFeatures);
    if (++i % batch_count == 0) batches.push_back(gen.GeneratorAlgoData());
  }
  if (gen.HasFeatures()) batches.push_back(gen.GeneratorAlgoData());
  auto columns = [](const std::map<std::string, std::vector<module::Value>>& m) {
    std::string o = "{";
    bool first = true;
    for (auto& kv : m) {
      if (!first) o += ",";
      first = false;
      o += Json::quote(kv.first) + ":[";
      for (size_t k = 0; k < kv.second.size(); ++k) { if (k) o += ","; o += json_value(kv.second[k]); }
      o += "]";
    }
    return o + "}";
  };
  std::string o = "[";
  for (size_t b = 0; b < batches.size(); ++b) {
    const easyrec::PBRequest& r = batches[b].Request;
    if (b) o += ",";
    o += "{\"user_features\":{";
    bool first = true;
    for (auto& kv : r.UserFeatures) { if (!first) o += ","; first = false; o += Json::quote(kv.first) + ":" + json_value(kv.second); }
    o += "},\"item_ids\":[";
    for (size_t k = 0; k < r.ItemIds.size(); ++k) { if (k) o += ","; o += Json::quote(r.ItemIds[k]); }
    o += "],\"context_features\":" + columns(r.ContextFeatures) + ",\"item_features\":" + columns(r.ItemFeatures) + "}";
  }
  o += "]";
  const long long need = (long long)o.size() + 1;
  if (out && cap >= (unsigned long long)need) memcpy(out, o.c_str(), (size_t)need);
  return need;
}

// An embedding hook (sort.RegisterEmbeddingHook, sort/dpp_sort.go:56-58) backed by a host table: item id -> row of
// emb[n_ids][dim] (ids without a row get zeros).  Stands for a user-registered Go function in the tests.
int ph_register_embedding_hook(ph_server* s, const char* hook_name, const char* const* ids, const double* emb,
                               unsigned long long n_ids, int dim) {
  if (!s || !hook_name || (!ids && n_ids) || !emb || dim <= 0) { g_err = "null argument"; return 1; }
  auto table = std::make_shared<std::unordered_map<std::string, std::vector<double>>>();
  for (unsigned long long i = 0; i < n_ids; ++i) (*table)[ids[i]] = std::vector<double>(emb + i * dim, emb + (i + 1) * dim);
  sort::RegisterEmbeddingHook(hook_name, [table, dim](context::RecommendContext*, const module::ItemPtr& it) {
    auto f = table->find(it->Id);
    return f == table->end() ? std::vector<double>((size_t)dim, 0.0) : f->second;
  });
  return 0;
}
int ph_set_user_vector(ph_server* s, const char* uid, const char* vector_string) {
  if (!s || !uid || !vector_string) { g_err = "null argument"; return 1; }
  s->vectors->m[uid] = vector_string;
  return 0;
}

// POST /api/recommend body -> response body.  Returns the number of bytes needed (incl. NUL); writes when it fits.
long long ph_recommend(ph_server* s, const char* request_json, char* out, unsigned long long cap) {
  if (!s || !request_json) { g_err = "null argument"; return -1; }
  Json req;
  std::string err;
  if (!Json::parse(request_json, &req, &err)) { g_err = "bad request JSON: " + err; return -1; }
  context::RecommendContext ctx;
  ctx.Config = &s->conf;
  ctx.RecommendId = "req-" + std::to_string(s->next_id++);
  ctx.Size = req["size"].as_int(10);  // Default_Size (recommend_controller.go:20-22)
  if (ctx.Size <= 0) ctx.Size = 10;
  ctx.Debug = req["debug"].as_bool(false);
  ctx.Param["scene"] = req["scene_id"].as_string();
  ctx.Param["category"] = req["category"].as_string().empty() ? "default" : req["category"].as_string();
  module::User user;
  user.Id = req["uid"].as_string();
  for (auto& kv : req["features"].obj) {
    if (kv.second.type == Json::Number) user.Properties[kv.first] = kv.second.num;
    else if (kv.second.type == Json::String) user.Properties[kv.first] = kv.second.str;
  }
  auto items = service::Recommend(&user, &ctx);
  std::string o = "{\"request_id\":" + Json::quote(ctx.RecommendId);
  if ((int)items.size() < ctx.Size) o += ",\"code\":299,\"msg\":\"items size not enough\"";
  else o += ",\"code\":200,\"msg\":\"success\"";
  o += ",\"size\":" + std::to_string(items.size()) + ",\"items\":[";
  for (size_t i = 0; i < items.size(); ++i) {
    if (i) o += ",";
    o += "{\"item_id\":" + Json::quote(items[i]->Id) + ",\"item_type\":" + Json::quote(items[i]->ItemType) +
         ",\"retrieve_id\":" + Json::quote(items[i]->RetrieveId) + ",\"score\":" + Json::number(items[i]->Score) + "}";
  }
  o += "],\"log\":[";
  for (size_t i = 0; i < ctx.Log.size(); ++i) { if (i) o += ","; o += Json::quote(ctx.Log[i]); }
  o += "]}";
  const long long need = (long long)o.size() + 1;
  if (out && cap >= (unsigned long long)need) memcpy(out, o.c_str(), (size_t)need);
  return need;
}

// doSort's head (DoSortHead / EmbeddingMissAboveThreshold) as a pure function, for the CPU tests: scores[n] and
// has_emb[n] in, the indices of the list the reference holds when it loads the embeddings out (returns their count);
// *missed = 1 when the share of items without an embedding exceeds the threshold (doSort then returns that list).
long long ph_dosort_head(const double* scores, const unsigned char* has_emb, int n, int size, int candidate_cnt,
                         double min_score_percent, double miss_threshold, int always_sort, int* out_idx, int* missed) {
  if (!scores || !has_emb || !out_idx || !missed || n < 0) { g_err = "null argument"; return -1; }
  std::vector<module::ItemPtr> items;
  for (int i = 0; i < n; ++i) { auto it = module::NewItem(std::to_string(i)); it->Score = scores[i]; items.push_back(it); }
  auto head = sort::DoSortHead(items, size, candidate_cnt, min_score_percent, always_sort != 0);
  size_t missing = 0;
  for (size_t i = 0; i < head.size(); ++i) {
    out_idx[i] = std::stoi(head[i]->Id);
    missing += has_emb[out_idx[i]] == 0;
  }
  *missed = sort::EmbeddingMissAboveThreshold(missing, head.size(), miss_threshold > 0 ? miss_threshold : 0.5) ? 1 : 0;
  return (long long)head.size();
}

// Item-feature column sets -> id-encoded fields (ingest::FieldEncoder).  spec_json: [{"column": "...", "vocab": [...]} |
// {"column": "...", "id": true}, ...]; rows_json: one object per item = the properties a FeatureDao fetched for it
// (a NULL column is simply absent).  out: n_items x n_fields u32.  Returns n_items * n_fields, or -1.
long long ph_encode_fields(const char* spec_json, const char* rows_json, unsigned int* out, unsigned long long cap) {
  if (!spec_json || !rows_json) { g_err = "null argument"; return -1; }
  Json spec, rows;
  std::string err;
  if (!Json::parse(spec_json, &spec, &err) || spec.type != Json::Array) { g_err = "bad field spec JSON: " + err; return -1; }
  if (!Json::parse(rows_json, &rows, &err) || rows.type != Json::Array) { g_err = "bad rows JSON: " + err; return -1; }
  std::vector<ingest::FieldSpec> specs;
  for (const Json& e : spec.arr) {
    ingest::FieldSpec fs;
    fs.Column = e["column"].as_string();
    fs.IsId = e["id"].as_bool(false);
    for (const Json& v : e["vocab"].arr) {
      if (v.type == Json::String) fs.Vocab.push_back(v.str);
      else if (v.type == Json::Number)
        fs.Vocab.push_back(ingest::ToString(v.num == std::floor(v.num) && std::fabs(v.num) < 9e15 ? module::Value((int64_t)v.num)
                                                                                                : module::Value(v.num)));
    }
    specs.push_back(std::move(fs));
  }
  const ingest::FieldEncoder enc(std::move(specs));
  const size_t F = enc.size(), N = rows.arr.size();
  if (out && cap >= N * F) {
    for (size_t i = 0; i < N; ++i) {
      module::Features props;
      for (auto& kv : rows.arr[i].obj) {
        if (kv.second.type == Json::String) props[kv.first] = kv.second.str;
        else if (kv.second.type == Json::Number) {
          if (kv.second.num == std::floor(kv.second.num) && std::fabs(kv.second.num) < 9e15) props[kv.first] = (int64_t)kv.second.num;
          else props[kv.first] = kv.second.num;
        }   // null / other kinds: no property, like a NULL column
      }
      enc.Encode(props, out + i * F);
    }
  }
  return (long long)(N * F);
}

// one "v" of the user-vector text (vector_recall.go:78-79): strconv.ParseFloat(v, 32), error ignored
float ph_parse_float32(const char* text) { return ingest::ParseFloat32(text ? text : ""); }

// sort/dpp_sort.go:224-233 embedding text -> doubles; returns the element count (writes up to cap)
long long ph_parse_embedding(const char* text, const char* sep, double* out, unsigned long long cap) {
  auto v = ingest::ParseEmbeddingText(text ? text : "", sep ? sep : "");
  for (size_t i = 0; i < v.size() && i < cap; ++i) out[i] = v[i];
  return (long long)v.size();
}
// recall cache round trip (vector_recall.go:35-58,103-110): ids/scores -> string -> JSON list of {item_id, score}
long long ph_recall_cache_roundtrip(const char* const* ids, const double* scores, int n, const char* model, char* out,
                                    unsigned long long cap, char* cache_out, unsigned long long cache_cap) {
  std::vector<module::ItemPtr> items;
  for (int i = 0; i < n; ++i) { auto it = module::NewItem(ids[i]); it->Score = scores[i]; items.push_back(it); }
  const std::string cache = ingest::FormatRecallCache(items, model);
  if (cache_out && cache_cap > cache.size()) memcpy(cache_out, cache.c_str(), cache.size() + 1);
  auto back = ingest::ParseRecallCache(cache, model, "");
  std::string o = "[";
  for (size_t i = 0; i < back.size(); ++i) {
    if (i) o += ",";
    o += "{\"item_id\":" + Json::quote(back[i]->Id) + ",\"score\":" + Json::number(back[i]->Score) + ",\"retrieve_id\":" +
         Json::quote(back[i]->RetrieveId) + "}";
  }
  o += "]";
  if (out && cap > o.size()) memcpy(out, o.c_str(), o.size() + 1);
  return (long long)o.size() + 1;
}

// alinkFMResponseFunc + GetScore (algorithm/eas/fm_response.go:28-53): response body in, one score per entry out.
// Returns the number of entries (writes up to cap), or -1 with ph_last_error set.
long long ph_alink_fm_scores(const char* body, double* out, unsigned long long cap) {
  algorithm::AlgoResponses res;
  Error e = algorithm::eas::AlinkFMResponseFunc(body ? body : "", &res);
  if (!e.empty()) { g_err = e; return -1; }
  for (size_t i = 0; i < res.size() && i < cap; ++i) out[i] = res[i]->GetScore();
  return (long long)res.size();
}

// tfservingResponseFunc (algorithm/tfserving/response.go:51-63): outputs [rows][cols] -> one score per value, row-major
long long ph_tfserving_scores(const double* outputs, int rows, int cols, double* out, unsigned long long cap) {
  if (rows < 0 || cols < 0 || (!outputs && rows * cols > 0)) { g_err = "bad arguments"; return -1; }
  std::vector<std::vector<double>> o((size_t)rows);
  for (int r = 0; r < rows; ++r) o[(size_t)r].assign(outputs + (size_t)r * cols, outputs + (size_t)(r + 1) * cols);
  auto res = algorithm::tfserving::TfservingResponseFunc(o);
  for (size_t i = 0; i < res.size() && i < cap; ++i) out[i] = res[i]->GetScore();
  return (long long)res.size();
}

// utils/ast known-answer entry: evaluates an expression over named values (names[i] -> values[i])
int ph_eval_expr(const char* expr, const char* const* names, const double* values, int n, double* out) {
  std::shared_ptr<ast::Expr> e;
  Error er = ast::Parse(expr ? expr : "", &e);
  if (!er.empty()) { g_err = er; return 1; }
  er = ast::Eval(*e, [&](const std::string& nm, double* o) {
    for (int i = 0; i < n; ++i)
      if (nm == names[i]) { *o = values[i]; return true; }
    return false;
  }, out);
  if (!er.empty()) { g_err = er; return 1; }
  return 0;
}

}  // extern "C"
