// pairec_host.cpp — implementation of the host-side mirror declared in pairec_host.hpp.
#include "pairec_host.hpp"

#include <algorithm>
#include <cctype>
#include <charconv>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <sstream>

namespace pairec {

// ================================================================================================ recconf
namespace recconf {
static std::vector<std::string> str_list(const Json& j) {
  std::vector<std::string> v;
  if (j.type == Json::Array)
    for (auto& e : j.arr)
      if (e.type == Json::String) v.push_back(e.str);
  return v;
}
Error LoadConfig(const std::string& json, RecommendConfig* out) {
  Json root;
  std::string err;
  if (!Json::parse(json, &root, &err)) return "invalid recconf JSON: " + err;
  if (root.type != Json::Object) return "recconf JSON must be an object";
  *out = RecommendConfig();
  out->RunMode = root["RunMode"].as_string();
  for (auto& a : root["AlgoConfs"].arr) {
    AlgoConfig c;
    c.Name = a["Name"].as_string();
    c.Type = a["Type"].as_string();
    c.LookupConf.FieldName = a["LookupConf"]["FieldName"].as_string();
    out->AlgoConfs.push_back(c);
  }
  for (auto& r : root["RecallConfs"].arr) {
    RecallConfig c;
    c.Name = r["Name"].as_string();
    c.RecallType = r["RecallType"].as_string();
    c.RecallAlgo = r["RecallAlgo"].as_string();
    c.ItemType = r["ItemType"].as_string();
    c.RecallCount = r["RecallCount"].as_int();
    out->RecallConfs.push_back(c);
  }
  for (auto& sc : root["SceneConfs"].obj)
    for (auto& cat : sc.second.obj) out->SceneConfs[sc.first][cat.first].RecallNames = str_list(cat.second["RecallNames"]);
  for (auto& rc : root["RankConf"].obj) {
    RankConfig c;
    c.RankAlgoList = str_list(rc.second["RankAlgoList"]);
    c.RankScore = rc.second["RankScore"].as_string();
    c.Processor = rc.second["Processor"].as_string();
    c.ContextFeatures = str_list(rc.second["ContextFeatures"]);
    c.ItemFeatures = str_list(rc.second["ItemFeatures"]);
    c.BatchCount = rc.second["BatchCount"].as_int();
    out->RankConf[rc.first] = c;
  }
  for (auto& s : root["SortNames"].obj) out->SortNames[s.first] = str_list(s.second);
  for (auto& s : root["FilterNames"].obj) out->FilterNames[s.first] = str_list(s.second);
  for (auto& s : root["SortConfs"].arr) {
    SortConfig c;
    c.Name = s["Name"].as_string();
    c.SortType = s["SortType"].as_string();
    c.SortByField = s["SortByField"].as_string();
    c.SwitchThreshold = s["SwitchThreshold"].as_number();
    const Json& d = s["DPPConf"];
    c.DPPConf.Name = d["Name"].as_string();
    c.DPPConf.Alpha = d["Alpha"].as_number();
    c.DPPConf.WindowSize = d["WindowSize"].as_int();
    c.DPPConf.AbortRunCount = d["AbortRunCount"].as_int();
    c.DPPConf.CandidateCount = d["CandidateCount"].as_int();
    c.DPPConf.MinScorePercent = d["MinScorePercent"].as_number();
    c.DPPConf.EmbMissedThreshold = d["EmbMissedThreshold"].as_number();
    c.DPPConf.NormalizeEmb = d["NormalizeEmb"].as_string();
    c.DPPConf.EnsurePositiveSim = d["EnsurePositiveSim"].as_string();
    c.DPPConf.TableName = d["TableName"].as_string();
    for (auto& v : d["EmbeddingHookNames"].arr) c.DPPConf.EmbeddingHookNames.push_back(v.as_string());
    c.DPPConf.FilterRetrieveIds = str_list(d["FilterRetrieveIds"]);
    const Json& sd = s["SSDConf"];
    c.SSDConf.Name = sd["Name"].as_string();
    c.SSDConf.Gamma = sd["Gamma"].as_number();
    c.SSDConf.UseSSDStar = sd["UseSSDStar"].as_bool(false);
    c.SSDConf.WindowSize = sd["WindowSize"].as_int();
    c.SSDConf.AbortRunCount = sd["AbortRunCount"].as_int();
    c.SSDConf.CandidateCount = sd["CandidateCount"].as_int();
    c.SSDConf.MinScorePercent = sd["MinScorePercent"].as_number();
    c.SSDConf.EmbMissedThreshold = sd["EmbMissedThreshold"].as_number();
    c.SSDConf.NormalizeEmb = sd["NormalizeEmb"].as_string();
    c.SSDConf.FilterRetrieveIds = str_list(sd["FilterRetrieveIds"]);
    out->SortConfs.push_back(c);
  }
  for (auto& f : root["FilterConfs"].arr) {
    FilterConfig c;
    c.Name = f["Name"].as_string();
    c.FilterType = f["FilterType"].as_string();
    c.RetainNum = f["RetainNum"].as_int();
    c.ShuffleItem = f["ShuffleItem"].as_bool(false);
    out->FilterConfs.push_back(c);
  }
  for (auto& g : root["GeneralRankConfs"].obj) {
    GeneralRankConfig c;
    const Json& rc = g.second["RankConf"];
    c.RankConf.RankAlgoList = str_list(rc["RankAlgoList"]);
    c.RankConf.RankScore = rc["RankScore"].as_string();
    c.RankConf.Processor = rc["Processor"].as_string();
    c.RankConf.ContextFeatures = str_list(rc["ContextFeatures"]);
    c.RankConf.ItemFeatures = str_list(rc["ItemFeatures"]);
    c.RankConf.BatchCount = rc["BatchCount"].as_int();
    for (auto& a : g.second["ActionConfs"].arr) c.ActionConfs.push_back({a["ActionType"].as_string(), a["ActionName"].as_string()});
    out->GeneralRankConfs[g.first] = c;
  }
  out->UserDefineConfs = root["UserDefineConfs"];
  return "";
}
}  // namespace recconf

// ================================================================================================ module
namespace module {
double ToFloat(const Value& v, double def) {
  if (auto p = std::get_if<double>(&v)) return *p;
  if (auto p = std::get_if<int64_t>(&v)) return (double)*p;
  if (auto p = std::get_if<std::string>(&v)) {
    char* end = nullptr;
    const double d = strtod(p->c_str(), &end);
    return (end && *end == 0 && end != p->c_str()) ? d : def;
  }
  return def;
}
ItemPtr NewItem(const std::string& id) { return std::make_shared<Item>(id); }
void Item::AddProperty(const std::string& k, Value v) {
  std::lock_guard<std::mutex> g(mutex);
  Properties[k] = std::move(v);
}
void Item::AddAlgoScore(const std::string& name, double score) {
  std::lock_guard<std::mutex> g(mutex);
  algoScores[name] = score;
}
Features Item::GetFeatures() {
  std::lock_guard<std::mutex> g(mutex);
  if (!RetrieveId.empty() && !Properties.count(RetrieveId)) {  // module/item.go:233-239
    Properties[RetrieveId] = Score;
    Properties["recall_name"] = RetrieveId;
    Properties["recall_score"] = Score;
  }
  return Properties;
}
Error Item::FloatExprData(const std::string& name, double* out) {
  std::lock_guard<std::mutex> g(mutex);
  if (name == "current_score") {  // :190-198
    algoScores["recall_score"] = Score;
    *out = Score;
    return "";
  }
  auto it = algoScores.find(name);
  if (it != algoScores.end()) { *out = it->second; return ""; }
  auto pt = Properties.find(name);
  if (pt != Properties.end()) { *out = ToFloat(pt->second, 0); return ""; }
  *out = 0;
  return "not found,name:" + name;
}
Features User::MakeUserFeatures() const {   // module/user.go:137-159
  Features features;
  for (auto& kv : Properties) {
    if (kv.first == "type") continue;
    if (auto str = std::get_if<std::string>(&kv.second)) {   // strconv.ParseFloat(str, 64) succeeds -> float64
      const char* b = str->c_str();
      char* e = nullptr;
      const double v = std::strtod(b, &e);
      // (ParseFloat takes the whole string or fails; it accepts no leading / trailing blanks)
      if (!str->empty() && e == b + str->size() && !std::isspace((unsigned char)b[0])) { features[kv.first] = v; continue; }
    }
    features[kv.first] = kv.second;
  }
  return features;
}
}  // namespace module

// ================================================================================================ algorithm
namespace algorithm {
namespace {
struct ScoreResponse : response::AlgoResponse {
  double score;
  explicit ScoreResponse(double s) : score(s) {}
  double GetScore() const override { return score; }
};
}  // namespace

AlgorithmFactory& Factory() {
  static AlgorithmFactory f;
  return f;
}
void AlgorithmFactory::Init(const std::vector<recconf::AlgoConfig>& confs) {
  std::unique_lock<std::shared_mutex> g(mutex_);
  for (auto& conf : confs) {
    std::shared_ptr<IAlgorithm> algo;
    if (conf.Type == "LOOKUP") algo = std::make_shared<LookupPolicy>();
    else continue;  // EAS / FAISS / TFSERVING / SELDON are the remote backends this repo replaces; GPU algos register in code
    if (algo->Init(&conf).empty()) algorithms_[conf.Name] = algo;
  }
}
Error AlgorithmFactory::Run(const std::string& name, const AlgoData& data, AlgoResult* out) {
  std::shared_ptr<IAlgorithm> algo;
  {
    std::shared_lock<std::shared_mutex> g(mutex_);
    auto it = algorithms_.find(name);
    if (it != algorithms_.end()) algo = it->second;
  }
  if (!algo) return "not found algorithm, name:" + name;  // algorithm/algorithm.go:113
  return algo->Run(data, out);
}
void AlgorithmFactory::RegisterAlgorithm(const std::string& name, std::shared_ptr<IAlgorithm> a) {
  std::unique_lock<std::shared_mutex> g(mutex_);
  algorithms_[name] = std::move(a);
}

Error eas::AlinkFMResponseFunc(const std::string& body, AlgoResponses* out) {
  Json j;
  std::string err;
  if (!Json::parse(body, &j, &err) || j.type != Json::Array)   // (bodyFormat, fm_response.go:55-61)
    return "error:" + (err.empty() ? std::string("not a JSON list") : err) + ", body:" + body.substr(0, 512);
  out->clear();
  for (const Json& e : j.arr) {
    auto r = std::make_shared<AlinkFMResponse>();
    r->Result = e["prediction_result"].as_number();
    r->Score = e["prediction_score"].as_number();
    out->push_back(std::move(r));
  }
  return "";
}

AlgoResponses tfserving::TfservingResponseFunc(const std::vector<std::vector<double>>& outputs) {
  AlgoResponses ret;
  for (auto& val : outputs)
    for (double score : val) ret.push_back(std::make_shared<ScoreResponse>(score));
  return ret;
}

Error LookupPolicy::Init(const recconf::AlgoConfig* conf) {
  conf_ = conf->LookupConf;
  return "";
}
Error LookupPolicy::Run(const AlgoData& algoData, AlgoResult* out) {
  auto pp = std::get_if<const FeatureList*>(&algoData);
  if (!pp || !*pp) return "LookupPolicy: algoData is not []map[string]interface{}";
  const FeatureList& list = **pp;
  if (list.empty()) { *out = std::monostate{}; return ""; }  // algorithm/lookup.go:39-41 returns nil, nil
  AlgoResponses res(list.size());
  for (size_t i = 0; i < list.size(); ++i) {
    auto it = list[i].find(conf_.FieldName);
    if (it == list[i].end()) { res[i] = std::make_shared<ScoreResponse>(0.5); continue; }  // :47-49
    auto d = std::get_if<double>(&it->second);
    if (!d) return "LookupPolicy: field " + conf_.FieldName + " is not float64";  // the reference panics (:45)
    res[i] = std::make_shared<ScoreResponse>(*d);
  }
  *out = std::move(res);
  return "";
}

Error GpuVectorAlgorithm::Run(const AlgoData& algoData, AlgoResult* out) {
  auto pp = std::get_if<const pai_web::VectorRequest*>(&algoData);
  if (!pp || !*pp) return "GpuVectorAlgorithm: algoData is not *pai_web.VectorRequest";
  const pai_web::VectorRequest& req = **pp;
  if (req.K == 0 || req.Vector.empty()) return "GpuVectorAlgorithm: empty request";
  // the C ABI takes no vector length; vector_recall.go:72-82 silently skips malformed "i:v" pairs, so short vectors occur
  if (req.Vector.size() != (size_t)prg_item_dim(cat_->h))
    return "GpuVectorAlgorithm: user vector has " + std::to_string(req.Vector.size()) + " elements, the item matrix " +
           std::to_string(prg_item_dim(cat_->h));
  const int k = (int)req.K;
  std::vector<uint32_t> rows((size_t)k);
  std::vector<float> scores((size_t)k);
  int32_t n = 0;
  if (prg_recall_topk(cat_->h, req.Vector.data(), 1, k, rows.data(), scores.data(), &n, PRG_MEM_HOST) != PRG_OK)
    return std::string("prg_recall_topk: ") + prg_last_error();
  pai_web::VectorReply reply;
  for (int i = 0; i < n; ++i) {
    if (rows[i] >= cat_->ids.size()) continue;
    reply.Retval.push_back(rows[i]);
    reply.Labels.push_back(cat_->ids[rows[i]]);
    reply.Scores.push_back(scores[i]);
  }
  *out = std::move(reply);
  return "";
}

Error GpuRankAlgorithm::Run(const AlgoData& algoData, AlgoResult* out) {
  auto pp = std::get_if<const FeatureList*>(&algoData);
  if (!pp || !*pp) return "GpuRankAlgorithm: algoData is not []map[string]interface{}";
  const FeatureList& list = **pp;
  if (list.empty()) { *out = std::monostate{}; return ""; }
  std::vector<uint32_t> rows(list.size(), 0xFFFFFFFFu);
  for (size_t i = 0; i < list.size(); ++i) {
    auto it = list[i].find("item_id");
    if (it == list[i].end()) continue;
    if (auto s = std::get_if<std::string>(&it->second)) {
      auto r = cat_->row_of.find(*s);
      if (r != cat_->row_of.end()) rows[i] = r->second;
    }
  }
  std::vector<double> sc(list.size());
  if (prg_rank(cat_->h, model_, rows.data(), 1, (int)rows.size(), sc.data(), PRG_MEM_HOST) != PRG_OK)
    return std::string("prg_rank: ") + prg_last_error();
  AlgoResponses res(list.size());
  for (size_t i = 0; i < list.size(); ++i) res[i] = std::make_shared<ScoreResponse>(sc[i]);
  *out = std::move(res);
  return "";
}
}  // namespace algorithm

void GpuCatalog::SetIds(std::vector<std::string> v) {
  ids = std::move(v);
  row_of.clear();
  row_of.reserve(ids.size() * 2);
  for (size_t i = 0; i < ids.size(); ++i) row_of.emplace(ids[i], (uint32_t)i);
}

// ================================================================================================ recall
namespace recall {
namespace {
std::map<std::string, std::shared_ptr<Recall>>& registry() {
  static std::map<std::string, std::shared_ptr<Recall>> m;
  return m;
}
// VectorDao reading the "i:v i:v ..." string from a user property (stands in for the Redis/Hologres DAOs)
struct PropertyVectorDao : VectorDao {
  std::function<bool(const std::string&, std::string*)> fn;
  Error VectorString(const std::string& id, std::string* out) override { return fn && fn(id, out) ? "" : VectoryEmptyError; }
};
}  // namespace
const Error VectoryEmptyError = "vector empty";
void RegisterRecall(const std::string& name, std::shared_ptr<Recall> r) { registry()[name] = std::move(r); }
std::shared_ptr<Recall> GetRecall(const std::string& name) {
  auto it = registry().find(name);
  return it == registry().end() ? nullptr : it->second;
}
void Load(const recconf::RecommendConfig&) {}  // config-selectable recall types are registered by the embedding program

VectorRecall::VectorRecall(const recconf::RecallConfig& conf, std::shared_ptr<VectorDao> dao)
    : modelName_(conf.Name), itemType_(conf.ItemType), recallAlgo_(conf.RecallAlgo), recallCount_(conf.RecallCount),
      dao_(std::move(dao)) {}

std::vector<module::ItemPtr> VectorRecall::GetCandidateItems(module::User* user, context::RecommendContext* ctx) {
  std::vector<module::ItemPtr> ret;
  std::string value;
  Error err = dao_->VectorString(user->Id, &value);  // vector_recall.go:59
  if (!err.empty()) {
    if (err != VectoryEmptyError) ctx->LogError("module=VectorRecall\tname=" + modelName_ + "\terr=" + err);
    return ret;
  }
  pai_web::VectorRequest request;
  request.K = (uint32_t)recallCount_;
  std::istringstream ss(value);  // "idx:val idx:val" (:70-82)
  std::string vc;
  while (std::getline(ss, vc, ' ')) {
    const size_t c = vc.find(':');
    if (c == std::string::npos || vc.find(':', c + 1) != std::string::npos) continue;
    request.Vector.push_back(ingest::ParseFloat32(vc.substr(c + 1)));
  }
  if (request.Vector.empty()) {
    ctx->LogError("module=VectorRecall\terror=user Vector empty");
    return ret;
  }
  algorithm::AlgoResult result;
  err = algorithm::Run(recallAlgo_, &request, &result);  // :88
  if (!err.empty()) {
    ctx->LogError("module=VectorRecall\terror=" + err);
    return ret;
  }
  auto reply = std::get_if<pai_web::VectorReply>(&result);
  if (!reply) return ret;
  for (size_t i = 0; i < reply->Labels.size(); ++i) {  // :93-102
    auto item = module::NewItem(reply->Labels[i]);
    item->RetrieveId = modelName_;
    item->ItemType = itemType_;
    item->Score = (double)reply->Scores[i];
    ret.push_back(item);
  }
  return ret;
}

std::vector<module::ItemPtr> ContextItemRecall::GetCandidateItems(module::User*, context::RecommendContext*) {
  std::vector<module::ItemPtr> out;
  out.reserve(items_.size());
  for (auto& it : items_) {  // fresh Items per request, like the reference's recalls
    auto n = module::NewItem(it->Id);
    n->Score = it->Score;
    n->RetrieveId = name_;
    n->ItemType = it->ItemType;
    n->Properties = it->Properties;
    out.push_back(n);
  }
  return out;
}
}  // namespace recall

// ================================================================================================ ast (utils/ast)
namespace ast {
struct Expr {
  enum Kind { Num, Param, Bin } kind = Num;
  double val = 0;
  std::string name;  // Param name or operator
  std::shared_ptr<Expr> lhs, rhs;
};
namespace {
struct Tok { std::string tok; int type; };  // 0 literal, 1 operator, 2 parameter (utils/ast/parse.go)
Error tokenize(const std::string& s, std::vector<Tok>* out) {
  size_t i = 0;
  while (i < s.size()) {
    const char c = s[i];
    if (c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r') { ++i; continue; }
    if (strchr("#()+-*/^%", c)) { out->push_back({std::string(1, c), 1}); ++i; continue; }
    if (c >= '0' && c <= '9') {
      const size_t st = i;
      while (i < s.size() && ((s[i] >= '0' && s[i] <= '9') || s[i] == '.' || s[i] == '_' || s[i] == 'e')) ++i;
      std::string t = s.substr(st, i - st);
      t.erase(std::remove(t.begin(), t.end(), '_'), t.end());
      out->push_back({t, 0});
      continue;
    }
    if (c == '$' && i + 1 < s.size() && s[i + 1] == '{') {
      const size_t e = s.find('}', i);
      if (e == std::string::npos) return "unterminated ${";
      out->push_back({s.substr(i + 2, e - i - 2), 2});
      i = e + 1;
      continue;
    }
    return std::string("symbol error: unkown '") + c + "'";
  }
  return "";
}
int prec(const std::string& op) {
  if (op == "+" || op == "-") return 20;
  if (op == "*" || op == "/" || op == "%") return 40;
  if (op == "^") return 60;
  if (op == "#") return 80;
  return -1;
}
struct Parser {
  const std::vector<Tok>& t;
  size_t i = 0;
  Error err;
  const Tok* cur() const { return i < t.size() ? &t[i] : nullptr; }
  std::shared_ptr<Expr> primary() {
    const Tok* c = cur();
    if (!c) return nullptr;
    if (c->type == 0 || (c->type == 1 && c->tok != "(")) {
      if (c->type == 1) { err = "want '(' or '0-9' but get '" + c->tok + "'"; return nullptr; }
      auto e = std::make_shared<Expr>();
      e->kind = Expr::Num;
      e->val = strtod(c->tok.c_str(), nullptr);
      ++i;
      return e;
    }
    if (c->type == 2) {
      auto e = std::make_shared<Expr>();
      e->kind = Expr::Param;
      e->name = c->tok;
      ++i;
      return e;
    }
    ++i;  // "("
    auto e = expression();
    if (!e) return nullptr;
    if (!cur() || cur()->tok != ")") { err = "want ')'"; return nullptr; }
    ++i;
    return e;
  }
  int cur_prec() const { return (cur() && cur()->type == 1) ? prec(cur()->tok) : -1; }
  std::shared_ptr<Expr> binop_rhs(int exec_prec, std::shared_ptr<Expr> lhs) {  // ast.go:167-196
    for (;;) {
      const int tok_prec = cur_prec();
      if (tok_prec < exec_prec) return lhs;
      const std::string op = cur()->tok;
      ++i;
      if (!cur()) return lhs;
      auto rhs = primary();
      if (!rhs) return nullptr;
      if (tok_prec < cur_prec()) {
        rhs = binop_rhs(tok_prec + 1, rhs);
        if (!rhs) return nullptr;
      }
      auto b = std::make_shared<Expr>();
      b->kind = Expr::Bin;
      b->name = op;
      b->lhs = lhs;
      b->rhs = rhs;
      lhs = b;
    }
  }
  std::shared_ptr<Expr> expression() {
    auto l = primary();
    if (!l) return nullptr;
    return binop_rhs(0, l);
  }
};
}  // namespace
Error Parse(const std::string& src, std::shared_ptr<Expr>* out) {
  std::vector<Tok> toks;
  Error e = tokenize(src, &toks);
  if (!e.empty()) return e;
  if (toks.empty()) return "empty token";
  Parser p{toks, 0, ""};
  auto r = p.expression();
  if (!r) return p.err.empty() ? "parse error" : p.err;
  *out = r;
  return "";
}
Error Eval(const Expr& e, const std::function<bool(const std::string&, double*)>& param, double* out) {
  switch (e.kind) {
    case Expr::Num: *out = e.val; return "";
    case Expr::Param: if (!param(e.name, out)) *out = 0.0; return "";  // ast.go:256-262 (unknown -> 0)
    case Expr::Bin: {
      double l = 0, r = 0;
      Error er = Eval(*e.lhs, param, &l);
      if (!er.empty()) return er;
      er = Eval(*e.rhs, param, &r);
      if (!er.empty()) return er;
      const std::string& op = e.name;
      if (op == "#") *out = (l != 0.0) ? l : r;
      else if (op == "^") *out = std::pow(l, r);
      else if (op == "+") *out = l + r;
      else if (op == "-") *out = l - r;
      else if (op == "*") *out = l * r;
      else if (op == "/") {
        if (r == 0) return "violation of arithmetic specification: a division by zero in ExprASTResult";  // panics upstream
        *out = l / r;
      } else if (op == "%") {
        // float64(int(l) % int(r)); Go's int() of a NaN / out-of-range float64 is the minimum int64 on amd64
        auto to_int = [](double v) -> long long {
          return (v >= -9223372036854775808.0 && v < 9223372036854775808.0) ? (long long)v : (long long)0x8000000000000000ull;
        };
        const long long li = to_int(l), ri = to_int(r);
        if (ri == 0) return "integer divide by zero";
        *out = ri == -1 ? 0.0 : (double)(li % ri);   // (x % -1 is 0; the machine instruction traps on MinInt64 % -1)
      } else *out = 0;
      return "";
    }
  }
  return "";
}
}  // namespace ast

// ================================================================================================ rank
namespace rank {
void Rank(module::User* user, std::vector<module::ItemPtr>& items, context::RecommendContext* ctx) {
  const std::string scene = ctx->GetParameter("scene");
  if (!ctx->Config) return;
  auto rc = ctx->Config->RankConf.find(scene);
  if (rc == ctx->Config->RankConf.end()) return;  // rank_service.go:153-157: no config -> no rank
  RankWithConfig(rc->second, user, items, ctx);
}
namespace {
std::map<std::string, std::vector<std::shared_ptr<IRank>>>& rank_inters() {
  static std::map<std::string, std::vector<std::shared_ptr<IRank>>> m;
  return m;
}
}  // namespace
void RegisterRank(const std::string& scene, std::shared_ptr<IRank> r) { rank_inters()[scene].push_back(std::move(r)); }
void ResetRanks() { rank_inters().clear(); }

void RankWithConfig(const recconf::RankConfig& rankConfig, module::User* user, std::vector<module::ItemPtr>& all_items,
                    context::RecommendContext* ctx) {
  int batchCount = rankConfig.BatchCount > 0 ? rankConfig.BatchCount : 100;  // :163-166
  if (rankConfig.RankAlgoList.empty() && rankConfig.RankScore.empty()) return;  // :168-171 (before any custom rank runs)
  // Processor "EasyRec" (eas.Eas_Processor_EASYREC, algorithm/eas/model.go:25): MakeUserFeatures2 and the columnar
  // generator (rank_service.go:173-182)
  const bool easyrecProcessor = rankConfig.Processor == "EasyRec";
  const module::Features userFeatures =
      !user ? module::Features() : (easyrecProcessor ? user->MakeUserFeatures2() : user->MakeUserFeatures());
  // custom ranks (rank_service.go:131-137, :185-204): the first IRank whose Filter claims an item takes it, with
  // features = item.GetFeatures() overlaid by the user's (custom_rank.go:33-43); the rest goes to the algorithms
  std::vector<module::ItemPtr> items;
  {
    auto ri = rank_inters().find(ctx->GetParameter("scene"));
    struct Claimed { std::vector<module::ItemPtr> items; algorithm::FeatureList data; };
    std::vector<Claimed> claimed(ri == rank_inters().end() ? 0 : ri->second.size());
    for (auto& it : all_items) {
      bool taken = false;
      for (size_t r = 0; r < claimed.size() && !taken; ++r) {
        if (!ri->second[r]->Filter(user, it, ctx)) continue;
        module::Features f = it->GetFeatures();
        for (auto& kv : userFeatures) f[kv.first] = kv.second;
        claimed[r].data.push_back(std::move(f));
        claimed[r].items.push_back(it);
        taken = true;
      }
      if (!taken) items.push_back(it);
    }
    for (size_t r = 0; r < claimed.size(); ++r)   // :237-246 (one goroutine per custom rank upstream)
      if (!claimed[r].items.empty()) ri->second[r]->Rank(user, claimed[r].items, claimed[r].data, ctx);
    if (items.empty()) return;   // :248-253
  }
  std::shared_ptr<ast::Expr> exprAst;
  if (!rankConfig.RankScore.empty()) {
    Error e = ast::Parse(rankConfig.RankScore, &exprAst);
    if (!e.empty()) { ctx->LogError("module=rank\trankscore=" + rankConfig.RankScore + "\terror=" + e); exprAst = nullptr; }
  }
  // one generator per request (its column layout is fixed by the config or by the first item); it is emptied per batch
  std::unique_ptr<EasyrecAlgoDataGenerator> egen;
  if (easyrecProcessor) {
    egen = std::make_unique<EasyrecAlgoDataGenerator>(rankConfig.ContextFeatures);
    egen->SetItemFeatures(rankConfig.ItemFeatures);
  }
  const bool wantItemFeatures = !rankConfig.ContextFeatures.empty() || !rankConfig.ItemFeatures.empty();   // :207-212
  for (size_t b0 = 0; b0 < items.size(); b0 += (size_t)batchCount) {
    const size_t b1 = std::min(items.size(), b0 + (size_t)batchCount);
    algorithm::FeatureList feats;
    EasyrecAlgoDataGenerator::AlgoData edata;
    if (egen) {
      for (size_t i = b0; i < b1; ++i) {
        module::Features f;
        if (wantItemFeatures) f = items[i]->GetFeatures();
        egen->AddFeatures(items[i], wantItemFeatures ? &f : nullptr, userFeatures);
      }
      edata = egen->GeneratorAlgoData();
    } else {
      feats.reserve(b1 - b0);
      for (size_t i = b0; i < b1; ++i) {  // AlgoDataGenerator.AddFeatures (algo_data.go:104-118): user ∪ item features
        module::Features f = userFeatures;
        for (auto& kv : items[i]->GetFeatures()) f[kv.first] = kv.second;
        feats.push_back(std::move(f));
      }
    }
    for (auto& algoName : rankConfig.RankAlgoList) {  // :264-289 (one goroutine per batch x algo upstream)
      algorithm::AlgoResult result;
      Error e = egen ? algorithm::Run(algoName, &edata.Request, &result) : algorithm::Run(algoName, &feats, &result);
      if (!e.empty()) { ctx->LogError("module=rank\terror=run algorithm error(" + e + ")"); continue; }  // :274-277
      auto res = std::get_if<algorithm::AlgoResponses>(&result);
      if (!res) continue;
      for (size_t j = 0; j < res->size() && b0 + j < b1; ++j) {  // :313-334
        auto& it = items[b0 + j];
        if ((*res)[j]->GetModuleType())
          for (auto& kv : (*res)[j]->GetScoreMap()) it->AddAlgoScore(algoName + "_" + kv.first, kv.second);
        else it->AddAlgoScore(algoName, (*res)[j]->GetScore());
      }
    }
    if (exprAst) {  // :339-363
      for (size_t i = b0; i < b1; ++i) {
        auto& it = items[i];
        double v = 0;
        Error e = ast::Eval(*exprAst, [&](const std::string& n, double* o) { return it->FloatExprData(n, o).empty(); }, &v);
        if (!e.empty()) { ctx->LogError("module=rank\terror=" + e); continue; }
        it->Score = v;
      }
    }
  }
}

GpuRank::GpuRank(std::shared_ptr<GpuCatalog> cat, std::string name, int model, std::vector<ingest::FieldSpec> user_fields,
                 std::vector<std::string> dense_columns, int heads)
    : cat_(std::move(cat)), name_(std::move(name)), model_(model), heads_(heads < 1 ? 1 : heads),
      user_enc_(std::make_shared<ingest::FieldEncoder>(std::move(user_fields))), dense_(std::move(dense_columns)) {}

void GpuRank::Rank(module::User* user, std::vector<module::ItemPtr>& items, const algorithm::FeatureList&,
                   context::RecommendContext* ctx) {
  if (items.empty()) return;
  std::vector<uint32_t> rows(items.size());
  for (size_t i = 0; i < items.size(); ++i) {
    auto r = cat_->row_of.find(items[i]->Id);
    rows[i] = r == cat_->row_of.end() ? 0xFFFFFFFFu : r->second;   // unknown id: padding row, score 0
  }
  const module::Features uf = user ? user->MakeUserFeatures() : module::Features();
  std::vector<uint32_t> uid(user_enc_->size());
  if (!uid.empty()) user_enc_->Encode(uf, uid.data());
  std::vector<float> dense(dense_.size(), 0.f);
  for (size_t c = 0; c < dense_.size(); ++c) {
    auto it = uf.find(dense_[c]);
    if (it == uf.end()) continue;
    if (auto d = std::get_if<double>(&it->second)) dense[c] = (float)*d;
    else if (auto i64 = std::get_if<int64_t>(&it->second)) dense[c] = (float)*i64;
  }
  prg_user_features u{uid.empty() ? nullptr : uid.data(), dense.empty() ? nullptr : dense.data()};
  std::vector<double> sc(items.size()), smap(heads_ > 1 ? items.size() * (size_t)heads_ : 0);
  if (prg_rank_ex(cat_->h, model_, rows.data(), 1, (int)rows.size(), &u, sc.data(), smap.empty() ? nullptr : smap.data(),
                  PRG_MEM_HOST) != PRG_OK) {
    ctx->LogError(std::string("module=rank\terror=run algorithm error(prg_rank_ex: ") + prg_last_error() + ")");  // :274-277
    return;   // scores stay as they are, as when algorithm.Run fails upstream
  }
  for (size_t i = 0; i < items.size(); ++i) {
    if (heads_ > 1)
      for (int o = 0; o < heads_; ++o) items[i]->AddAlgoScore(name_ + "_" + std::to_string(o), smap[i * (size_t)heads_ + o]);
    items[i]->AddAlgoScore(name_, sc[i]);
    items[i]->Score = sc[i];
  }
}

// ---- Processor "EasyRec": service/rank/algo_data.go:173-350
EasyrecAlgoDataGenerator::EasyrecAlgoDataGenerator(const std::vector<std::string>& contextFeatures) {
  for (auto& name : contextFeatures) itemFeatures_.push_back({name, 2});   // reflect.TypeOf(""): a missing value is ""
}
void EasyrecAlgoDataGenerator::SetItemFeatures(const std::vector<std::string>& inputItemFeatures) {
  if (!inputItemFeatures.empty()) {
    hasInputItemFeatureMap_ = true;
    if (inputItemFeatures[0] != "*") {
      parseInputItemFeature_ = true;
      for (auto& name : inputItemFeatures) inputItemFeatures_.push_back({name, 2});
    }   // "*": every feature of the first item that is not a context feature (AddFeatures)
  } else {
    parseInputItemFeature_ = true;
  }
}
module::Value EasyrecAlgoDataGenerator::DefaultValue(const Feature& f) {
  if (f.kind == 0) return 0.0;
  if (f.kind == 1) return (int64_t)0;
  return std::string();
}
void EasyrecAlgoDataGenerator::AddFeatures(const module::ItemPtr& item, const module::Features* itemFeatures,
                                           const module::Features& userFeatures) {
  static const module::Features kNone;
  const module::Features& feats = itemFeatures ? *itemFeatures : kNone;
  if (item) requestItem_.push_back(item);
  if (!parseFeature_) {   // (:243-254; the constructor above always sets parseFeature, as upstream's does)
    for (auto& kv : feats) itemFeatures_.push_back({kv.first, kv.second.index()});
    userFeatures_ = userFeatures;
    parseFeature_ = true;
  }
  if (!parseInputItemFeature_) {   // ItemFeatures == ["*"]: the first item decides the input item feature columns
    for (auto& kv : feats) {
      bool isContext = false;
      for (auto& cf : itemFeatures_) isContext |= cf.name == kv.first;
      if (!isContext) inputItemFeatures_.push_back({kv.first, kv.second.index()});
    }
    parseInputItemFeature_ = true;
  }
  if (userFeatures_.empty()) userFeatures_ = userFeatures;
  for (auto& f : itemFeatures_) {
    auto it = feats.find(f.name);
    contextFeatures_[f.name].push_back(it != feats.end() ? it->second : DefaultValue(f));
  }
  if (hasInputItemFeatureMap_)
    for (auto& f : inputItemFeatures_) {
      auto it = feats.find(f.name);
      inputItemFeatureMap_[f.name].push_back(it != feats.end() ? it->second : DefaultValue(f));
    }
}
EasyrecAlgoDataGenerator::AlgoData EasyrecAlgoDataGenerator::GeneratorAlgoData() {
  AlgoData d;
  d.Items = requestItem_;
  d.Request.UserFeatures = userFeatures_;                                  // builder.AddUserFeature per entry
  for (auto& it : requestItem_) d.Request.ItemIds.push_back(it->Id);       // builder.AddItemId
  for (auto& kv : contextFeatures_) { d.Request.ContextFeatures[kv.first] = kv.second; kv.second.clear(); }
  for (auto& kv : inputItemFeatureMap_) { d.Request.ItemFeatures[kv.first] = kv.second; kv.second.clear(); }
  requestItem_.clear();
  return d;
}
}  // namespace rank

namespace algorithm {
namespace {
struct EasyrecResponse : response::AlgoResponse {   // algorithm/eas/easyrec_response.go:13-33, multiValModule = true
  std::map<std::string, double> scoreArr;
  double GetScore() const override { return 0; }
  std::map<std::string, double> GetScoreMap() const override { return scoreArr; }
  bool GetModuleType() const override { return true; }
};
}  // namespace
GpuEasyrecAlgorithm::GpuEasyrecAlgorithm(std::shared_ptr<GpuCatalog> cat, int model, std::vector<ingest::FieldSpec> user_fields,
                                         std::vector<std::string> dense_columns, std::vector<std::string> outputs)
    : cat_(std::move(cat)), model_(model), user_enc_(std::make_shared<ingest::FieldEncoder>(std::move(user_fields))),
      dense_(std::move(dense_columns)), outputs_(std::move(outputs)) {}

Error GpuEasyrecAlgorithm::Run(const AlgoData& algoData, AlgoResult* out) {
  auto pp = std::get_if<const easyrec::PBRequest*>(&algoData);
  if (!pp || !*pp) return "GpuEasyrecAlgorithm: algoData is not *easyrec.PBRequest";
  const easyrec::PBRequest& req = **pp;
  if (req.ItemIds.empty()) { *out = std::monostate{}; return ""; }
  const size_t n = req.ItemIds.size(), heads = outputs_.size() > 1 ? outputs_.size() : 1;
  std::vector<uint32_t> rows(n);
  for (size_t i = 0; i < n; ++i) {
    auto r = cat_->row_of.find(req.ItemIds[i]);
    rows[i] = r == cat_->row_of.end() ? 0xFFFFFFFFu : r->second;   // no such item: zeros (easyrec_response.go:56-61)
  }
  std::vector<uint32_t> uid(user_enc_->size());
  if (!uid.empty()) user_enc_->Encode(req.UserFeatures, uid.data());
  std::vector<float> dense(dense_.size(), 0.f);
  for (size_t c = 0; c < dense_.size(); ++c) {
    auto it = req.UserFeatures.find(dense_[c]);
    if (it != req.UserFeatures.end()) dense[c] = (float)module::ToFloat(it->second, 0.0);
  }
  prg_user_features u{uid.empty() ? nullptr : uid.data(), dense.empty() ? nullptr : dense.data()};
  std::vector<double> sc(n), smap(heads > 1 ? n * heads : 0);
  if (prg_rank_ex(cat_->h, model_, rows.data(), 1, (int)n, &u, sc.data(), smap.empty() ? nullptr : smap.data(), PRG_MEM_HOST) != PRG_OK)
    return std::string("prg_rank_ex: ") + prg_last_error();
  AlgoResponses res(n);
  for (size_t i = 0; i < n; ++i) {
    if (heads == 1) { res[i] = std::make_shared<ScoreResponse>(sc[i]); continue; }
    auto r = std::make_shared<EasyrecResponse>();
    for (size_t o = 0; o < heads; ++o) r->scoreArr[outputs_[o]] = smap[i * heads + o];
    res[i] = std::move(r);
  }
  *out = std::move(res);
  return "";
}
}  // namespace algorithm

// ================================================================================================ feature
namespace feature {
namespace {
std::map<std::string, LoadFeatureFunc>& load_funcs() {
  static std::map<std::string, LoadFeatureFunc> m;
  return m;
}
}  // namespace
void RegisterLoadFeatureFunc(const std::string& scene, LoadFeatureFunc f) { load_funcs()[scene] = std::move(f); }
void LoadFeatures(module::User* user, std::vector<module::ItemPtr>& items, context::RecommendContext* ctx) {
  // feature_service.go:77-131.  The reference runs the scene's FeatureDaos first (replaced by the HBM tables) and the
  // LoadFeatureFunc beside them; upstream additionally requires a FeatureConfs entry for the scene (:87)
  auto it = load_funcs().find(ctx->GetParameter("scene"));
  if (it != load_funcs().end() && it->second) it->second(user, items, ctx);
}
LoadFeatureFunc ItemIdProperty() {
  return [](module::User*, std::vector<module::ItemPtr>& items, context::RecommendContext*) {
    for (auto& it : items) it->AddProperty("item_id", it->Id);
  };
}
void ResetLoadFuncs() { load_funcs().clear(); }
}  // namespace feature

// ================================================================================================ sort
namespace sort {
namespace {
std::map<std::string, EmbeddingHookFunc>& embedding_hooks() {
  static std::map<std::string, EmbeddingHookFunc> m;
  return m;
}
std::map<std::string, std::shared_ptr<ISort>>& mapping() {
  static std::map<std::string, std::shared_ptr<ISort>> m;
  return m;
}
std::map<std::string, std::vector<std::shared_ptr<ISort>>>& strategies() {
  static std::map<std::string, std::vector<std::shared_ptr<ISort>>> m;
  return m;
}
// sort.Sort(sort.Reverse(ItemScoreSlice(items))) with Go's pdqsort tie order (prg_sort_desc_host)
void go_sort_desc(std::vector<module::ItemPtr>& items, const std::vector<double>& key) {
  const int n = (int)items.size();
  if (n <= 1) return;
  std::vector<int32_t> perm((size_t)n);
  prg_sort_desc_host(key.data(), n, perm.data());
  std::vector<module::ItemPtr> out((size_t)n);
  for (int i = 0; i < n; ++i) out[(size_t)i] = items[(size_t)perm[(size_t)i]];
  items.swap(out);
}
}  // namespace
void RegisterSort(const std::string& name, std::shared_ptr<ISort> s) {
  if (!mapping().count(name)) mapping()[name] = std::move(s);
}
Error GetSort(const std::string& name, std::shared_ptr<ISort>* out) {
  auto it = mapping().find(name);
  if (it == mapping().end()) return "ISort not found, name:" + name;
  *out = it->second;
  return "";
}
void Load(const recconf::RecommendConfig& c) {
  RegisterSort("ItemRankScore", std::make_shared<ItemRankScoreSort>());  // sort/item_rank_score.go:34-36 init()
  for (auto& sc : c.SortConfs)
    if (sc.SortType == "AlgoScoreSort") RegisterSort(sc.Name, std::make_shared<AlgoScoreSort>(sc));
  for (auto& kv : c.SortNames) {
    std::vector<std::shared_ptr<ISort>> v;
    for (auto& n : kv.second) {
      auto it = mapping().find(n);
      if (it != mapping().end()) v.push_back(it->second);
    }
    strategies()[kv.first] = v;
  }
}
void Sort(SortData* data, const std::string& tag) {
  context::RecommendContext* ctx = data->Context;
  std::string scene = ctx->GetParameter("scene") + tag;
  std::string category = ctx->GetParameter("category");
  if (category.empty()) category = "default";
  std::vector<std::shared_ptr<ISort>> sorts;
  auto it = strategies().find(scene);
  if (it == strategies().end()) it = strategies().find(category);
  if (it != strategies().end()) sorts = it->second;
  if (sorts.empty()) {
    sorts.push_back(std::make_shared<ItemRankScoreSort>());
    ctx->LogInfo("defaultSort=ItemRankScore\tscene=" + scene);
  }
  for (auto& s : sorts) s->Sort(data);  // the return value is ignored upstream (sort.go:123)
}

Error ItemRankScoreSort::Sort(SortData* d) {
  std::vector<double> key(d->Data.size());
  for (size_t i = 0; i < key.size(); ++i) key[i] = d->Data[i]->Score;
  go_sort_desc(d->Data, key);
  return "";
}
AlgoScoreSort::AlgoScoreSort(const recconf::SortConfig& c)
    : sortByField_(c.SortByField.empty() ? "current_score" : c.SortByField), switchThreshold_(c.SwitchThreshold) {}
Error AlgoScoreSort::Sort(SortData* d) {
  double maxScore = -1e300;  // GetMaxScore (:28-36)
  for (auto& it : d->Data) maxScore = std::max(maxScore, it->Score);
  const std::string field = maxScore > switchThreshold_ ? "current_score" : sortByField_;
  std::vector<double> key(d->Data.size());
  for (size_t i = 0; i < key.size(); ++i) {
    if (!d->Data[i]->FloatExprData(field, &key[i]).empty()) {
      key[i] = d->Data[i]->Score;
      d->Context->LogInfo("get sort field " + field + " from item " + d->Data[i]->Id + " failed");
    }
  }
  // sort.Slice(less = iScore > jScore) — same pdqsort.  Deliberate deviation on an error path: when item j has no such
  // field the reference's `less` overwrites iScore with items[j].Score and compares it with the error value of jScore
  // (algo_score_sort.go:57-61, a slip: not a strict weak order, so the resulting order depends on the comparison
  // sequence); here a failing item is keyed by its own Score, which is what the warning it logs says it does.
  go_sort_desc(d->Data, key);
  return "";
}

std::vector<module::ItemPtr> DoSortHead(std::vector<module::ItemPtr> items, int size, int candidateCnt,
                                        double minScorePercent, bool alwaysSort) {
  if (items.empty()) return items;
  const bool truncates = (candidateCnt > 0 || minScorePercent > 0) && (int)items.size() > size;
  if (alwaysSort || truncates) {   // ssd_sort.go:301 sorts first, always; dpp_sort.go:281 only when it truncates
    std::vector<double> key(items.size());
    for (size_t i = 0; i < items.size(); ++i) key[i] = items[i]->Score;
    go_sort_desc(items, key);
  }
  if (truncates) {
    if (candidateCnt > 0) {
      const int cnt = std::max(size, candidateCnt);
      if (cnt < (int)items.size()) items.resize((size_t)cnt);
    }
    if (minScorePercent > 0 && (int)items.size() > size) {
      int idx = size;
      const double maxScore = items[0]->Score;
      for (; idx < (int)items.size(); ++idx)
        if (items[(size_t)idx]->Score / maxScore < minScorePercent) break;
      items.resize((size_t)idx);
    }
  }
  return items;
}
bool EmbeddingMissAboveThreshold(size_t missing, size_t total, double threshold) {
  return total > 0 && (double)missing / (double)total > threshold;
}

void RegisterEmbeddingHook(const std::string& name, EmbeddingHookFunc fn) { embedding_hooks()[name] = std::move(fn); }

GpuDPPSort::GpuDPPSort(const recconf::DPPSortConfig& c, std::shared_ptr<GpuCatalog> cat) : conf_(c), cat_(std::move(cat)) {
  if (conf_.WindowSize <= 0) conf_.WindowSize = 10;  // dpp_sort.go:89-91
  if (conf_.EmbMissedThreshold <= 0) conf_.EmbMissedThreshold = 0.5;  // :85, :101-103
}
Error GpuDPPSort::Sort(SortData* d) {
  auto& candidates = d->Data;
  if (candidates.empty()) return "";
  context::RecommendContext* ctx = d->Context;
  if (conf_.AbortRunCount > 0 && (int)candidates.size() <= conf_.AbortRunCount) {  // :118-123
    ItemRankScoreSort().Sort(d);
    return "";
  }
  std::vector<module::ItemPtr> selected, backup;  // :145-158
  for (auto& it : candidates) {
    if (std::find(conf_.FilterRetrieveIds.begin(), conf_.FilterRetrieveIds.end(), it->RetrieveId) != conf_.FilterRetrieveIds.end())
      backup.push_back(it);
    else selected.push_back(it);
  }
  std::vector<module::ItemPtr> result = selected;
  if (!selected.empty()) {
    // what doSort holds when it loads the embeddings, and returns on every error path (:280-300, :304-307, :317-320):
    // already in Go's sort order and cut — the device gets exactly this list (no second presort / cut there)
    const std::vector<module::ItemPtr> head = DoSortHead(selected, ctx->Size, conf_.CandidateCount, conf_.MinScorePercent, false);
    // hooks (:352-370): hasHookFunc = any configured name is registered; hasTable = TableName set (an attached catalog
    // with a diversity table stands for it when the config names no hooks at all)
    std::vector<EmbeddingHookFunc> hooks;
    for (auto& name : conf_.EmbeddingHookNames) {
      auto h = embedding_hooks().find(name);
      if (h != embedding_hooks().end()) hooks.push_back(h->second);
    }
    const bool hasTable = !conf_.TableName.empty() || hooks.empty();
    std::vector<uint32_t> rows(head.size());
    std::vector<double> score(head.size());
    size_t missing = 0;
    for (size_t i = 0; i < head.size(); ++i) {
      auto r = cat_->row_of.find(head[i]->Id);
      rows[i] = r == cat_->row_of.end() ? 0xFFFFFFFEu : r->second;  // no embedding: a substitute direction (prg_dpp)
      missing += r == cat_->row_of.end();
      score[i] = head[i]->Score;
    }
    if (hasTable && EmbeddingMissAboveThreshold(missing, head.size(), conf_.EmbMissedThreshold)) {  // :246-249
      ctx->LogError("load embedding table cache failed the number of items missing embedding is above threshold");
      result = head;
      result.insert(result.end(), backup.begin(), backup.end());
      candidates.swap(result);
      return "";
    }
    std::vector<double> hook;
    int hook_dim = 0;
    bool hook_error = false;
    if (!hooks.empty()) {
      for (size_t i = 0; i < head.size() && !hook_error; ++i) {
        std::vector<double> e;                                       // GenerateEmbedding (:362-370): concatenation
        for (auto& fn : hooks) { auto part = fn(ctx, head[i]); e.insert(e.end(), part.begin(), part.end()); }
        if (i == 0) { hook_dim = (int)e.size(); hook.reserve(head.size() * e.size()); }
        if ((int)e.size() != hook_dim || hook_dim == 0) hook_error = true;   // :423-425 / :433-435
        hook.insert(hook.end(), e.begin(), e.end());
      }
    }
    prg_dpp_params p{};
    p.alpha = conf_.Alpha;
    p.top_n = ctx->Size;
    p.window_size = conf_.WindowSize;
    p.norm_mode = 0;
    p.normalize_emb = (conf_.NormalizeEmb == "false" || conf_.NormalizeEmb == "False") ? 0 : 1;
    p.no_positive_sim = (conf_.EnsurePositiveSim == "false" || conf_.EnsurePositiveSim == "False") ? 1 : 0;
    std::vector<int32_t> idx((size_t)std::max(1, ctx->Size), -1);
    int32_t n = 0, st = 0;
    int rc = PRG_OK;
    if (hook_error) {
      ctx->LogError("build kernel matrix failed the length of user-defined function is not equal");
      rc = PRG_EINVAL;
    } else {
      rc = prg_dpp_ex(cat_->h, rows.data(), score.data(), hook.empty() ? nullptr : hook.data(), hook_dim, hasTable ? 1 : 0, 1,
                      (int)rows.size(), &p, idx.data(), &n, &st, PRG_MEM_HOST);
      if (rc != PRG_OK) ctx->LogError(std::string("build kernel matrix failed ") + prg_last_error());  // :317-320: `return items`
    }
    if (rc != PRG_OK) {
      result = head;
    } else if (st != 0) {
      ctx->LogError("build kernel matrix failed all item score is zero");                 // :385-388, :397-400
      result = head;
    } else {
      for (size_t i = 0; i < head.size(); ++i) head[i]->AddAlgoScore("dpp_relevance_score", score[i]);  // :411: every item
      result.clear();
      for (int i = 0; i < n; ++i) result.push_back(head[(size_t)idx[(size_t)i]]);
    }
  }
  result.insert(result.end(), backup.begin(), backup.end());
  candidates.swap(result);
  return "";
}
GpuSSDSort::GpuSSDSort(const recconf::SSDSortConfig& c, std::shared_ptr<GpuCatalog> cat) : conf_(c), cat_(std::move(cat)) {
  if (conf_.Gamma <= 0) conf_.Gamma = 0.25;     // ssd_sort.go:62,81-83
  if (conf_.EmbMissedThreshold <= 0) conf_.EmbMissedThreshold = 0.5;  // :77, :96-98
  if (conf_.WindowSize <= 0) conf_.WindowSize = 5;  // :84-86
}
Error GpuSSDSort::Sort(SortData* d) {
  auto& candidates = d->Data;
  if (candidates.empty()) return "";
  context::RecommendContext* ctx = d->Context;
  if (conf_.AbortRunCount > 0 && (int)candidates.size() <= conf_.AbortRunCount) {  // :129-134
    ItemRankScoreSort().Sort(d);
    return "";
  }
  std::vector<module::ItemPtr> selected, backup;  // :158-176
  for (auto& it : candidates) {
    if (std::find(conf_.FilterRetrieveIds.begin(), conf_.FilterRetrieveIds.end(), it->RetrieveId) != conf_.FilterRetrieveIds.end())
      backup.push_back(it);
    else selected.push_back(it);
  }
  std::vector<module::ItemPtr> result = selected;
  if (!selected.empty()) {
    // doSort's list: sorted in Go's order and cut (ssd_sort.go:301-331); the device gets exactly this list
    const std::vector<module::ItemPtr> head = DoSortHead(selected, ctx->Size, conf_.CandidateCount, conf_.MinScorePercent, true);
    std::vector<uint32_t> rows(head.size());
    std::vector<double> score(head.size());
    size_t missing = 0;
    for (size_t i = 0; i < head.size(); ++i) {
      auto r = cat_->row_of.find(head[i]->Id);
      rows[i] = r == cat_->row_of.end() ? 0xFFFFFFFEu : r->second;
      missing += r == cat_->row_of.end();
      score[i] = head[i]->Score;
    }
    if (EmbeddingMissAboveThreshold(missing, head.size(), conf_.EmbMissedThreshold)) {  // ssd_sort.go:270-273, :334-337
      ctx->LogError("load embedding table cache failed the number of items missing embedding is above threshold");
      result = head;
      result.insert(result.end(), backup.begin(), backup.end());
      candidates.swap(result);
      return "";
    }
    prg_ssd_params p{};
    p.gamma = conf_.Gamma;
    p.top_n = ctx->Size;
    p.window_size = conf_.WindowSize;
    p.norm_mode = 0;
    p.normalize_emb = (conf_.NormalizeEmb == "false" || conf_.NormalizeEmb == "False") ? 0 : 1;
    p.use_ssd_star = conf_.UseSSDStar ? 1 : 0;
    std::vector<int32_t> idx((size_t)std::max(1, ctx->Size), -1);
    int32_t n = 0, st = 0;
    if (prg_ssd(cat_->h, rows.data(), score.data(), 1, (int)rows.size(), &p, idx.data(), &n, &st, PRG_MEM_HOST) != PRG_OK) {
      ctx->LogError(std::string("module=SSDSort\terror=") + prg_last_error());  // `return items`: the sorted, cut list
      result = head;
    } else {
      result.clear();
      for (int i = 0; i < n; ++i) result.push_back(head[(size_t)idx[(size_t)i]]);
    }
  }
  result.insert(result.end(), backup.begin(), backup.end());
  candidates.swap(result);
  return "";
}
}  // namespace sort

// ================================================================================================ filter
namespace filter {
void UniqueFilter(std::vector<module::ItemPtr>* items) {
  std::vector<module::ItemPtr> out;
  std::unordered_map<std::string, module::ItemPtr> uniq;
  for (auto& item : *items) {
    auto it = uniq.find(item->Id);
    if (it == uniq.end()) {
      uniq[item->Id] = item;
      out.push_back(item);
    } else {
      for (auto& kv : item->algoScores) it->second->AddAlgoScore(kv.first, kv.second);
      if (it->second->RecallScores.empty()) it->second->RecallScores[it->second->RetrieveId] = it->second->Score;
      it->second->RecallScores[item->RetrieveId] = item->Score;
    }
  }
  items->swap(out);
}
}  // namespace filter

namespace filter {
namespace {
std::map<std::string, std::shared_ptr<IFilter>>& fregistry() {
  static std::map<std::string, std::shared_ptr<IFilter>> m;
  return m;
}
struct UniqueFilterImpl : IFilter {
  Error Filter(FilterData* d) override { UniqueFilter(&d->Data); return ""; }
};
struct AdjustCountFilter : IFilter {  // filter/adjust_count_filter.go:37-77
  int retainNum;
  bool shuffleItem;
  Error Filter(FilterData* d) override {
    if ((int)d->Data.size() <= retainNum) return "";
    if (shuffleItem) return "AdjustCountFilter: ShuffleItem uses an unseeded math/rand upstream; not reproducible here";
    d->Data.resize((size_t)retainNum);
    return "";
  }
};
}  // namespace
void RegisterFilter(const std::string& name, std::shared_ptr<IFilter> f) { fregistry()[name] = std::move(f); }
Error GetFilter(const std::string& name, std::shared_ptr<IFilter>* out) {
  auto it = fregistry().find(name);
  if (it == fregistry().end()) return "Filter not found, name:" + name;
  *out = it->second;
  return "";
}
void Load(const recconf::RecommendConfig& c) {
  RegisterFilter("UniqueFilter", std::make_shared<UniqueFilterImpl>());
  for (auto& fc : c.FilterConfs) {
    if (fc.FilterType == "AdjustCountFilter") {
      auto f = std::make_shared<AdjustCountFilter>();
      f->retainNum = fc.RetainNum;
      f->shuffleItem = fc.ShuffleItem;
      RegisterFilter(fc.Name, f);
    }
  }
}
void ResetFilters() { fregistry().clear(); }
}  // namespace filter

// ================================================================================================ general rank
namespace general_rank {
std::vector<module::ItemPtr> Rank(module::User* user, std::vector<module::ItemPtr> items, context::RecommendContext* ctx) {
  if (!ctx->Config) return items;
  auto gc = ctx->Config->GeneralRankConfs.find(ctx->GetParameter("scene"));
  if (gc == ctx->Config->GeneralRankConfs.end()) return items;  // general_rank.go:216-262: no config, items pass through
  recconf::RankConfig rc = gc->second.RankConf;
  if (rc.BatchCount <= 0) rc.BatchCount = 100;                  // base_general_rank.go:58-60
  if (!rc.RankAlgoList.empty()) rank::RankWithConfig(rc, user, items, ctx);
  for (auto& ac : gc->second.ActionConfs) {                     // action.go:40-83
    if (ac.ActionType == "sort") {
      std::shared_ptr<sort::ISort> s;
      Error e = sort::GetSort(ac.ActionName, &s);
      if (!e.empty()) { ctx->LogError("create action error:" + e); continue; }
      sort::SortData sd;
      sd.Data = items; sd.Context = ctx; sd.User = user;
      s->Sort(&sd);
      items = sd.Data;
    } else if (ac.ActionType == "filter") {
      std::shared_ptr<filter::IFilter> f;
      Error e = filter::GetFilter(ac.ActionName, &f);
      if (!e.empty()) { ctx->LogError("create action error:" + e); continue; }
      filter::FilterData fd;
      fd.Data = items; fd.Uid = user ? user->Id : ""; fd.Context = ctx;
      e = f->Filter(&fd);
      if (!e.empty()) ctx->LogError("module=general_rank\tfilter error=" + e);
      items = fd.Data;
    } else {
      ctx->LogError("create action error:error to find actionType:" + ac.ActionType);
    }
  }
  return items;
}
}  // namespace general_rank

// ================================================================================================ ingest
namespace ingest {
std::vector<double> ParseEmbeddingText(const std::string& text, const std::string& sep_in) {
  const std::string sep = sep_in.empty() ? "," : sep_in;  // dpp_sort.go:92-94
  size_t a = 0, b = text.size();
  while (a < b && (text[a] == '{' || text[a] == '}')) ++a;   // strings.Trim(s, "{}")
  while (b > a && (text[b - 1] == '{' || text[b - 1] == '}')) --b;
  std::vector<double> out;
  const std::string body = text.substr(a, b - a);
  size_t pos = 0;
  for (;;) {
    const size_t nx = body.find(sep, pos);
    const std::string tok = body.substr(pos, nx == std::string::npos ? std::string::npos : nx - pos);
    char* end = nullptr;
    const double v = strtod(tok.c_str(), &end);
    out.push_back((end && end != tok.c_str() && *end == 0) ? v : 0.0);  // ParseFloat error -> element stays 0
    if (nx == std::string::npos) break;
    pos = nx + sep.size();
  }
  return out;
}
float ParseFloat32(const std::string& s) {
  // strconv.ParseFloat(s, 32): ONE rounding, decimal -> nearest float32 (strtof; going through a double first rounds
  // twice and can land on the other neighbour); the whole string or nothing — on any error the reference's ignored
  // error leaves 0 (vector_recall.go:78-79)
  if (s.empty() || std::isspace((unsigned char)s[0])) return 0.f;
  char* end = nullptr;
  const float v = std::strtof(s.c_str(), &end);
  return end == s.c_str() + s.size() ? v : 0.f;
}
std::string ToString(const module::Value& v) {
  if (const auto* s = std::get_if<std::string>(&v)) return *s;
  if (const auto* i = std::get_if<int64_t>(&v)) return std::to_string(*i);
  char buf[400];   // 'f' format of a double needs up to 309 integer digits
  const auto r = std::to_chars(buf, buf + sizeof buf, std::get<double>(v), std::chars_format::fixed);
  return std::string(buf, r.ptr);
}
FieldEncoder::FieldEncoder(std::vector<FieldSpec> specs) : specs_(std::move(specs)), rows_(specs_.size()) {
  for (size_t f = 0; f < specs_.size(); ++f)
    for (size_t i = 0; i < specs_[f].Vocab.size(); ++i) rows_[f].emplace(specs_[f].Vocab[i], (uint32_t)i);  // first wins
}
void FieldEncoder::Encode(const module::Features& properties, uint32_t* out) const {
  for (size_t f = 0; f < specs_.size(); ++f) {
    out[f] = kAbsent;
    const auto it = properties.find(specs_[f].Column);
    if (it == properties.end()) continue;   // NULL column: ParseColumnValues returned nil, no property was written
    if (specs_[f].IsId) {
      int64_t id = -1;
      if (const auto* i = std::get_if<int64_t>(&it->second)) id = *i;
      else if (const auto* d = std::get_if<double>(&it->second)) { if (*d >= 0 && *d == std::floor(*d) && *d < 4294967295.0) id = (int64_t)*d; }
      else { char* end = nullptr; const std::string& s = std::get<std::string>(it->second); const long long v = strtoll(s.c_str(), &end, 10); if (!s.empty() && end && *end == 0) id = v; }
      if (id >= 0 && id < 0xFFFFFFFFll) out[f] = (uint32_t)id;
    } else {
      const auto r = rows_[f].find(ToString(it->second));
      if (r != rows_[f].end()) out[f] = r->second;
    }
  }
}
// fmt's %v of a float64 = strconv.FormatFloat(v, 'g', -1, 64): the shortest digits that round-trip; %e form when the
// decimal exponent is < -4 or >= 6 (strconv/ftoa.go: "if precision was the shortest possible, use precision 6 for this
// decision"), at least two exponent digits; otherwise plain decimals.  C's %g decides with the digit count instead
// (100.0 would print as 1e+02).
std::string FormatFloatV(double v) {
  if (std::isnan(v)) return "NaN";
  if (std::isinf(v)) return v > 0 ? "+Inf" : "-Inf";
  char buf[64];
  const auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);   // shortest: d[.ddd]e[+-]XX
  std::string sci(buf, r.ptr);
  const size_t e = sci.find('e');
  const int exp = atoi(sci.c_str() + e + 1);
  if (exp < -4 || exp >= 6) return sci;   // (to_chars writes the exponent the way strconv does: sign, >= 2 digits)
  const auto f = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);           // shortest, no exponent
  return std::string(buf, f.ptr);
}
std::string FormatRecallCache(const std::vector<module::ItemPtr>& items, const std::string& modelName) {
  std::string s;
  for (size_t i = 0; i < items.size(); ++i) {
    if (i) s += ",";
    // fmt.Sprintf("%s:%s:%v", id, modelName, score) (vector_recall.go:106)
    s += items[i]->Id + ":" + modelName + ":" + FormatFloatV(items[i]->Score);
  }
  return s;
}
std::vector<module::ItemPtr> ParseRecallCache(const std::string& s, const std::string& modelName, const std::string& itemType) {
  std::vector<module::ItemPtr> ret;
  size_t pos = 0;
  while (pos <= s.size()) {
    const size_t nx = s.find(',', pos);
    const std::string id = s.substr(pos, nx == std::string::npos ? std::string::npos : nx - pos);
    module::ItemPtr item;
    if (id.find(':') != std::string::npos) {  // vector_recall.go:42-48: vars[0] = id, vars[2] = score
      std::vector<std::string> vars;
      size_t p = 0;
      for (;;) {
        const size_t c = id.find(':', p);
        vars.push_back(id.substr(p, c == std::string::npos ? std::string::npos : c - p));
        if (c == std::string::npos) break;
        p = c + 1;
      }
      item = module::NewItem(vars[0]);
      if (vars.size() > 2) item->Score = strtod(vars[2].c_str(), nullptr);
    } else {
      item = module::NewItem(id);
    }
    item->RetrieveId = modelName;
    item->ItemType = itemType;
    ret.push_back(item);
    if (nx == std::string::npos) break;
    pos = nx + 1;
  }
  return ret;
}
}  // namespace ingest

// ================================================================================================ service
namespace service {
std::vector<module::ItemPtr> Recommend(module::User* user, context::RecommendContext* ctx) {
  std::vector<module::ItemPtr> items;
  const recconf::RecommendConfig* conf = ctx->Config;
  if (!conf) return items;
  const std::string scene = ctx->GetParameter("scene");
  std::string category = ctx->GetParameter("category");
  if (category.empty()) category = "default";
  // RecallService.GetItems (service/recall.go:53-151): scene -> category -> recall names
  auto sc = conf->SceneConfs.find(scene);
  if (sc != conf->SceneConfs.end()) {
    auto cat = sc->second.find(category);
    if (cat == sc->second.end()) cat = sc->second.find("default");
    if (cat != sc->second.end()) {
      for (auto& name : cat->second.RecallNames) {
        auto r = recall::GetRecall(name);
        if (!r) { ctx->LogError("recall not found, name:" + name); continue; }
        auto got = r->GetCandidateItems(user, ctx);
        items.insert(items.end(), got.begin(), got.end());
      }
    }
  }
  // Filter (service/recommend.go:28-35): FilterNames[scene] or ["default"]
  auto fn = conf->FilterNames.find(scene);
  if (fn == conf->FilterNames.end()) fn = conf->FilterNames.find("default");
  if (fn != conf->FilterNames.end()) {
    for (auto& name : fn->second) {
      std::shared_ptr<filter::IFilter> f;
      if (!filter::GetFilter(name, &f).empty()) continue;
      filter::FilterData fd;
      fd.Data = items; fd.Uid = user ? user->Id : ""; fd.Context = ctx;
      f->Filter(&fd);
      items = fd.Data;
    }
  }
  items = general_rank::Rank(user, items, ctx);   // user_recommend.go:116
  feature::LoadFeatures(user, items, ctx);        // :129
  rank::Rank(user, items, ctx);                   // :137
  sort::SortData sd;
  sd.Data = items;
  sd.Context = ctx;
  sd.User = user;
  sort::Sort(&sd, "");
  items = sd.Data;
  if ((int)items.size() > ctx->Size) items.resize((size_t)ctx->Size);  // user_recommend.go:168
  return items;
}
}  // namespace service

namespace filter { void ResetFilters(); }
namespace feature { void ResetLoadFuncs(); }
void ResetRegistries() {
  filter::ResetFilters();
  feature::ResetLoadFuncs();
  rank::ResetRanks();
  sort::embedding_hooks().clear();
  sort::mapping().clear();
  sort::strategies().clear();
  recall::registry().clear();
  algorithm::Factory().~AlgorithmFactory();
  new (&algorithm::Factory()) algorithm::AlgorithmFactory();
}

}  // namespace pairec
