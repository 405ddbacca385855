// pairec_host.hpp — host-side mirror (C++17) of the reference's operator / plugin interfaces for the hot path.
//
// The reference host is Go; no Go toolchain exists in the build image or on the GPU box, so the layer that sits above
// the C ABI (include/pairec_gpu.h) is written in C++ with the SAME names, argument meaning and error behaviour as the
// Go interfaces it stands in for, so that tests/test_host_plugin*.py read like the reference's own plugin tests
// (construct Items + a RecommendContext, run the plugin, assert order/count):
//
//   recconf::RecommendConfig        recconf/recconf.go:48-93 (subset: AlgoConfs, RecallConfs, SceneConfs, RankConf,
//                                   SortNames, SortConfs{SortByField,SwitchThreshold,DPPConf}, FilterNames, UserDefineConfs)
//   module::Item / User             module/item.go:15-27,168-248 ; module/user.go
//   context::RecommendContext       context/recommend_context.go
//   algorithm::IAlgorithm, factory  algorithm/algorithm.go:28-31,107-168 ; LookupPolicy algorithm/lookup.go:37-51
//   pai_web::VectorRequest/Reply    algorithm/faiss/vectorretrieval.proto:11-20
//   recall::Recall, VectorRecall    service/recall/recall.go:18-34 ; service/recall/vector_recall.go:32-123
//   rank::RankService               service/rank/rank_service.go:102-372 (generic and EasyRec processors), utils/ast/ast.go:215-269
//   rank::EasyrecAlgoDataGenerator  service/rank/algo_data.go:173-350 (columnar easyrec.PBRequest: user features, item ids, columns)
//   sort::ISort, SortService        sort/sort.go:27-35,65-150 ; ItemRankScoreSort sort/item_rank_score.go:26-32 ;
//                                   AlgoScoreSort sort/algo_score_sort.go:38-66 ; DPPSort sort/dpp_sort.go:108-167
//   filter::UniqueFilter            filter/unique_filter.go:26-49
//   service::UserRecommendService   service/user_recommend.go:46-183 (recall -> filter -> rank -> sort -> truncate)
//
// GPU-backed plugins (GpuVectorAlgorithm, GpuRankAlgorithm, GpuEasyrecAlgorithm, rank::GpuRank, GpuDPPSort, GpuSSDSort)
// call libpairec_gpu.so through its C ABI only.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <variant>
#include <vector>

#include "../../include/pairec_gpu.h"
#include "json.hpp"

namespace pairec {

using Error = std::string;  // "" == nil

// ------------------------------------------------------------------------------------------------ recconf
namespace recconf {
struct LookupConfig { std::string FieldName; };
struct AlgoConfig { std::string Name, Type; LookupConfig LookupConf; };
struct RecallConfig { std::string Name, RecallType, RecallAlgo, ItemType; int RecallCount = 0; };
struct RankConfig {   // recconf.go:736-745
  std::vector<std::string> RankAlgoList, ContextFeatures, ItemFeatures;
  std::string RankScore, Processor;
  int BatchCount = 0;
};
struct DPPSortConfig {
  std::string Name, NormalizeEmb, EnsurePositiveSim, TableName;
  std::vector<std::string> EmbeddingHookNames;
  double Alpha = 0, MinScorePercent = 0, EmbMissedThreshold = 0;   // recconf.go:960-979
  int WindowSize = 0, AbortRunCount = 0, CandidateCount = 0;
  std::vector<std::string> FilterRetrieveIds;
};
struct SSDSortConfig {  // recconf/recconf.go:980-1000
  std::string Name, NormalizeEmb;
  double Gamma = 0, MinScorePercent = 0, EmbMissedThreshold = 0;
  bool UseSSDStar = false;
  int WindowSize = 0, AbortRunCount = 0, CandidateCount = 0;
  std::vector<std::string> FilterRetrieveIds;
};
struct SortConfig { std::string Name, SortType, SortByField; double SwitchThreshold = 0; DPPSortConfig DPPConf; SSDSortConfig SSDConf; };
struct SceneCategory { std::vector<std::string> RecallNames; };
struct FilterConfig { std::string Name, FilterType; int RetainNum = 0; bool ShuffleItem = false; };   // recconf.go FilterConfig
struct ActionConfig { std::string ActionType, ActionName; };                                          // :746-749
struct GeneralRankConfig { RankConfig RankConf; std::vector<ActionConfig> ActionConfs; };             // :934-938
struct RecommendConfig {
  std::string RunMode;
  std::vector<AlgoConfig> AlgoConfs;
  std::vector<RecallConfig> RecallConfs;
  std::map<std::string, std::map<std::string, SceneCategory>> SceneConfs;
  std::map<std::string, RankConfig> RankConf;
  std::map<std::string, std::vector<std::string>> SortNames, FilterNames;
  std::vector<SortConfig> SortConfs;
  std::vector<FilterConfig> FilterConfs;
  std::map<std::string, GeneralRankConfig> GeneralRankConfs;
  Json UserDefineConfs;
};
// recconf.LoadConfig (recconf/recconf.go:1089-1100) from a JSON string
Error LoadConfig(const std::string& json, RecommendConfig* out);
}  // namespace recconf

// ------------------------------------------------------------------------------------------------ module
namespace module {
using Value = std::variant<double, int64_t, std::string>;
using Features = std::map<std::string, Value>;
double ToFloat(const Value& v, double def);  // utils.ToFloat

struct Item {
  std::string Id;
  double Score = 0;
  std::string RetrieveId, ItemType;
  std::vector<double> Embedding;
  Features Properties;
  std::map<std::string, double> algoScores;
  std::map<std::string, double> RecallScores;
  mutable std::mutex mutex;

  explicit Item(std::string id) : Id(std::move(id)) {}
  void AddProperty(const std::string& k, Value v);
  void AddAlgoScore(const std::string& name, double score);                  // module/item.go:168-176
  Features GetFeatures();                                                     // :229-248 (injects recall_name/score)
  Error FloatExprData(const std::string& name, double* out);                  // :189-212
};
using ItemPtr = std::shared_ptr<Item>;
ItemPtr NewItem(const std::string& id);

struct User {
  std::string Id;
  Features Properties;
  Features MakeUserFeatures() const;                          // module/user.go:137-159: no "type", numeric strings -> float64
  Features MakeUserFeatures2() const { return Properties; }   // :161-167 (EasyRec processor): a clone
};
}  // namespace module

// ------------------------------------------------------------------------------------------------ context
namespace context {
struct RecommendContext {
  std::string RecommendId;
  int Size = 10;
  bool Debug = false;
  std::map<std::string, std::string> Param;  // "scene", "category", ...
  const recconf::RecommendConfig* Config = nullptr;
  std::vector<std::string> Log;               // LogInfo / LogError sink (the reference writes to glog)
  std::string GetParameter(const std::string& k) const { auto it = Param.find(k); return it == Param.end() ? "" : it->second; }
  void LogInfo(const std::string& m) { Log.push_back("INFO " + m); }
  void LogError(const std::string& m) { Log.push_back("ERROR " + m); }
};
}  // namespace context

// ------------------------------------------------------------------------------------------------ algorithm
namespace pai_web {
struct VectorRequest { uint32_t K = 0; std::vector<float> Vector; };
struct VectorReply { std::vector<uint64_t> Retval; std::vector<float> Scores; std::vector<std::string> Labels; };
}  // namespace pai_web

namespace easyrec {   // algorithm/eas/easyrec/easyrec_predict.proto:150-212 — the fields the rank path fills
using PBFeature = module::Value;   // the scalar kinds of the oneof (int / long / float / double / string feature)
struct PBRequest {
  std::map<std::string, PBFeature> UserFeatures;                    // user_features = 2
  std::vector<std::string> ItemIds;                                 // item_ids = 3
  std::map<std::string, std::vector<PBFeature>> ContextFeatures;    // context_features = 4: one value per item id
  std::map<std::string, std::vector<PBFeature>> ItemFeatures;       // item_features = 6
  std::map<std::string, std::string> MetaData;                      // meta_data = 7
};
}  // namespace easyrec

namespace algorithm {
namespace response {
struct AlgoResponse {  // algorithm/response/resonse.go:3-7
  virtual ~AlgoResponse() = default;
  virtual double GetScore() const = 0;
  virtual std::map<std::string, double> GetScoreMap() const { return {}; }
  virtual bool GetModuleType() const { return false; }
};
}  // namespace response
using FeatureList = std::vector<module::Features>;
using AlgoResponses = std::vector<std::shared_ptr<response::AlgoResponse>>;
// Go's interface{} payloads on this path: generic-processor feature maps, a faiss VectorRequest or (Processor "EasyRec")
// the columnar easyrec.PBRequest
using AlgoData = std::variant<std::monostate, const FeatureList*, const pai_web::VectorRequest*, const easyrec::PBRequest*>;
using AlgoResult = std::variant<std::monostate, AlgoResponses, pai_web::VectorReply>;

struct IAlgorithm {  // algorithm/algorithm.go:28-31
  virtual ~IAlgorithm() = default;
  virtual Error Init(const recconf::AlgoConfig* conf) = 0;
  virtual Error Run(const AlgoData& algoData, AlgoResult* out) = 0;
};
class AlgorithmFactory {
 public:
  void Init(const std::vector<recconf::AlgoConfig>& confs);                       // :48-67
  Error Run(const std::string& name, const AlgoData& data, AlgoResult* out);     // :107-120
  void RegisterAlgorithm(const std::string& name, std::shared_ptr<IAlgorithm> a);  // :164-168
 private:
  std::shared_mutex mutex_;
  std::map<std::string, std::shared_ptr<IAlgorithm>> algorithms_;
};
AlgorithmFactory& Factory();
inline void Load(const recconf::RecommendConfig& c) { Factory().Init(c.AlgoConfs); }
inline Error Run(const std::string& name, const AlgoData& d, AlgoResult* out) { return Factory().Run(name, d, out); }
inline void RegisterAlgorithm(const std::string& n, std::shared_ptr<IAlgorithm> a) { Factory().RegisterAlgorithm(n, std::move(a)); }

namespace eas {
// algorithm/eas/fm_response.go:13-53 — the ALINK_FM processor's wire form: per item the predicted LABEL
// ("prediction_result") and the probability of THAT label ("prediction_score"); GetScore() turns the pair back into
// P(label = 1): result == 0 ? 1 - score : score (:28-34).  The in-process FM (prg_rank, PRG_MODEL_FM) has no wire and
// emits P(1) directly; this restatement is for replaying recorded ALINK responses beside it.
struct AlinkFMResponse : response::AlgoResponse {
  double Result = 0, Score = 0;
  double GetScore() const override { return Result == 0.0 ? 1 - Score : Score; }
};
// alinkFMResponseFunc (:36-53): body = JSON list of {"prediction_result", "prediction_score"}; a body that does not parse
// is an error that quotes its first 512 bytes
Error AlinkFMResponseFunc(const std::string& body, AlgoResponses* out);
}  // namespace eas

namespace tfserving {
// tfservingResponseFunc (algorithm/tfserving/response.go:51-63): PredictResponse.Outputs [][]float64 — one row per item,
// one value per output — flattened row-major into one single-score response per VALUE.  prg_rank_ex's score map
// ([items][heads]) has the same layout, so a one-head tower gives one response per item in item order.
AlgoResponses TfservingResponseFunc(const std::vector<std::vector<double>>& outputs);
}  // namespace tfserving

class LookupPolicy : public IAlgorithm {  // algorithm/lookup.go
 public:
  Error Init(const recconf::AlgoConfig* conf) override;
  Error Run(const AlgoData& algoData, AlgoResult* out) override;
 private:
  recconf::LookupConfig conf_;
};
}  // namespace algorithm

// ------------------------------------------------------------------------------------------------ the GPU catalog
// Item identity in the reference is a string; the kernels work on u32 rows.  One catalog per engine: row i <-> ids[i].
struct GpuCatalog {
  prg_handle* h = nullptr;
  std::vector<std::string> ids;
  std::unordered_map<std::string, uint32_t> row_of;
  void SetIds(std::vector<std::string> v);
};

namespace algorithm {
// IAlgorithm behind a RecallAlgo name: VectorRequest -> VectorReply through prg_recall_topk (replaces algorithm/faiss)
class GpuVectorAlgorithm : public IAlgorithm {
 public:
  explicit GpuVectorAlgorithm(std::shared_ptr<GpuCatalog> c) : cat_(std::move(c)) {}
  Error Init(const recconf::AlgoConfig*) override { return ""; }
  Error Run(const AlgoData& algoData, AlgoResult* out) override;
 private:
  std::shared_ptr<GpuCatalog> cat_;
};
// IAlgorithm behind a RankAlgoList name: feature maps -> scores through prg_rank.  The stock feature maps
// (service/rank/algo_data.go:104-118 over module/item.go:229-248) carry NO item id: this adapter needs the "item_id"
// property that feature::ItemIdProperty() (a LoadFeatureFunc, service/feature/feature_service.go:36-38) writes; a map
// without it scores 0 (comma-ok lookup, never a panic).  It cannot see which request a batch belongs to, so user
// features do not reach the model through it — rank::GpuRank (the IRank plugin) is the full-featured drop-in.
class GpuRankAlgorithm : public IAlgorithm {
 public:
  GpuRankAlgorithm(std::shared_ptr<GpuCatalog> c, int model) : cat_(std::move(c)), model_(model) {}
  Error Init(const recconf::AlgoConfig*) override { return ""; }
  Error Run(const AlgoData& algoData, AlgoResult* out) override;
 private:
  std::shared_ptr<GpuCatalog> cat_;
  int model_;
};
}  // namespace algorithm

// ------------------------------------------------------------------------------------------------ recall
namespace recall {
struct Recall {  // service/recall/recall.go:18-20
  virtual ~Recall() = default;
  virtual std::vector<module::ItemPtr> GetCandidateItems(module::User* user, context::RecommendContext* ctx) = 0;
};
void RegisterRecall(const std::string& name, std::shared_ptr<Recall> r);  // :32-34
std::shared_ptr<Recall> GetRecall(const std::string& name);
void Load(const recconf::RecommendConfig& c);  // builds the config-selectable recalls it knows (VectorRecall)

struct VectorDao {  // module/vector_dao.go:13-15
  virtual ~VectorDao() = default;
  virtual Error VectorString(const std::string& id, std::string* out) = 0;
};
extern const Error VectoryEmptyError;
class VectorRecall : public Recall {  // service/recall/vector_recall.go
 public:
  VectorRecall(const recconf::RecallConfig& conf, std::shared_ptr<VectorDao> dao);
  std::vector<module::ItemPtr> GetCandidateItems(module::User* user, context::RecommendContext* ctx) override;
 private:
  std::string modelName_, itemType_, recallAlgo_;
  int recallCount_;
  std::shared_ptr<VectorDao> dao_;
};
// in-memory recall of a fixed item list (the role MockRecall / ContextItemRecall play for config 1)
class ContextItemRecall : public Recall {
 public:
  ContextItemRecall(std::string name, std::vector<module::ItemPtr> items) : name_(std::move(name)), items_(std::move(items)) {}
  std::vector<module::ItemPtr> GetCandidateItems(module::User*, context::RecommendContext*) override;
 private:
  std::string name_;
  std::vector<module::ItemPtr> items_;
};
}  // namespace recall

// ------------------------------------------------------------------------------------------------ rank
namespace ast {
// utils/ast: "${a} * 2 + ${b}" with ops # ^ + - * / % and parentheses; division by zero is an error (the reference panics)
struct Expr;
Error Parse(const std::string& src, std::shared_ptr<Expr>* out);
Error Eval(const Expr& e, const std::function<bool(const std::string&, double*)>& param, double* out);
}  // namespace ast
namespace ingest { struct FieldSpec; class FieldEncoder; }
namespace feature {
// service/feature/feature_service.go:20-38: the user-defined feature loader of a scene, run before rank
// (service/user_recommend.go:129).  The FeatureDaos themselves are replaced by the HBM-resident tables.
using LoadFeatureFunc = std::function<void(module::User*, std::vector<module::ItemPtr>&, context::RecommendContext*)>;
void RegisterLoadFeatureFunc(const std::string& scene, LoadFeatureFunc f);
void LoadFeatures(module::User* user, std::vector<module::ItemPtr>& items, context::RecommendContext* ctx);
LoadFeatureFunc ItemIdProperty();   // item.AddProperty("item_id", item.Id)
}
namespace rank {
// service/rank/custom_rank.go:8-13.  Items a custom rank claims (Filter) bypass the BatchCount-sized algorithm.Run
// fan-out and reach Rank() in ONE call per request, as Items, together with the User (rank_service.go:185-200,237-246)
struct IRank {
  virtual ~IRank() = default;
  virtual bool Filter(module::User* user, const module::ItemPtr& item, context::RecommendContext* ctx) = 0;
  virtual void Rank(module::User* user, std::vector<module::ItemPtr>& items, const algorithm::FeatureList& requestData,
                    context::RecommendContext* ctx) = 0;
};
void RegisterRank(const std::string& scene, std::shared_ptr<IRank> r);   // rank_service.go:45-60
void ResetRanks();
void Rank(module::User* user, std::vector<module::ItemPtr>& items, context::RecommendContext* ctx);  // RankService.Rank
// the shared core: batches -> algorithm.Run -> AddAlgoScore -> RankScore expression (rank_service.go:163-363)
void RankWithConfig(const recconf::RankConfig& rankConfig, module::User* user, std::vector<module::ItemPtr>& items,
                    context::RecommendContext* ctx);
}
namespace rank {
// The GPU rank as an IRank: every item of the request in one prg_rank_ex call, the user's categorical features encoded
// by `user_fields` (ingest::FieldEncoder over user.MakeUserFeatures()) and numeric context features by `dense_columns`.
// Writes algoScores[name] (and name_<head> for a multi-head tower) and Item.Score, as RankService does for its own
// algorithms (rank_service.go:313-363 with RankScore = "${name}").
class GpuRank : public IRank {
 public:
  GpuRank(std::shared_ptr<GpuCatalog> cat, std::string name, int model, std::vector<ingest::FieldSpec> user_fields,
          std::vector<std::string> dense_columns, int heads);
  bool Filter(module::User*, const module::ItemPtr&, context::RecommendContext*) override { return true; }
  void Rank(module::User* user, std::vector<module::ItemPtr>& items, const algorithm::FeatureList& requestData,
            context::RecommendContext* ctx) override;
 private:
  std::shared_ptr<GpuCatalog> cat_;
  std::string name_;
  int model_, heads_;
  std::shared_ptr<ingest::FieldEncoder> user_enc_;
  std::vector<std::string> dense_;
};
}
namespace rank {
// service/rank/algo_data.go:173-350 — the generator RankService uses when RankConf.Processor == "EasyRec"
// (CreateAlgoDataGenerator, :33-40): user features ONCE per request, the item ids, and one column per context / input item
// feature (a value per item, the feature's default where an item lacks it) instead of one merged map per item.
class EasyrecAlgoDataGenerator {
 public:
  struct AlgoData { std::vector<module::ItemPtr> Items; easyrec::PBRequest Request; };
  explicit EasyrecAlgoDataGenerator(const std::vector<std::string>& contextFeatures);      // :200-218
  void SetItemFeatures(const std::vector<std::string>& inputItemFeatures);                // :221-237 ("*": infer from the first item)
  // itemFeatures == nullptr: rank_service.go:207-212 fetches item.GetFeatures() only when the config names features
  void AddFeatures(const module::ItemPtr& item, const module::Features* itemFeatures, const module::Features& userFeatures);  // :239-290
  bool HasFeatures() const { return !requestItem_.empty(); }                               // :348-350
  AlgoData GeneratorAlgoData();                                                            // :292-325
 private:
  struct Feature { std::string name; size_t kind; };   // kind: alternative of module::Value (0 double, 1 int64, 2 string)
  static module::Value DefaultValue(const Feature& f);  // :159-176: 0 of the numeric type, "" otherwise
  std::vector<module::ItemPtr> requestItem_;
  std::map<std::string, std::vector<module::Value>> contextFeatures_, inputItemFeatureMap_;
  std::vector<Feature> itemFeatures_, inputItemFeatures_;
  module::Features userFeatures_;
  bool parseFeature_ = true, parseInputItemFeature_ = false, hasInputItemFeatureMap_ = false;
};
}
namespace algorithm {
// IAlgorithm behind a RankAlgoList name of a scene whose RankConf.Processor is "EasyRec": the stock RankService then
// hands it ONE easyrec.PBRequest per batch (rank_service.go:273 with algo_data.go:84-86) that carries the item ids and the
// request's user features — everything the device needs, with no plugin interface beyond IAlgorithm and no feature-load
// hook.  ItemIds -> rows through the catalog (an unknown id scores 0 on every output, easyrec_response.go:56-61),
// UserFeatures -> categorical user field ids (ingest::FieldEncoder) and dense context columns, then prg_rank_ex.  The
// request's context / item feature columns are not read: item features live in the HBM tables, keyed by row.
// One output: responses carry GetScore(); several (`outputs` names the tower's heads in order): multi-value responses
// (GetModuleType() == true, GetScoreMap() = {output: score}) as easyrecMutValResponseFunc builds them (:35-70), which
// RankService stores as algoScores[algo + "_" + output] (rank_service.go:313-334).
class GpuEasyrecAlgorithm : public IAlgorithm {
 public:
  GpuEasyrecAlgorithm(std::shared_ptr<GpuCatalog> cat, int model, std::vector<ingest::FieldSpec> user_fields,
                      std::vector<std::string> dense_columns, std::vector<std::string> outputs);
  Error Init(const recconf::AlgoConfig*) override { return ""; }
  Error Run(const AlgoData& algoData, AlgoResult* out) override;
 private:
  std::shared_ptr<GpuCatalog> cat_;
  int model_;
  std::shared_ptr<ingest::FieldEncoder> user_enc_;
  std::vector<std::string> dense_, outputs_;
};
}
namespace general_rank {
// GeneralRankService.Rank (service/general_rank/general_rank.go:216) -> BaseGeneralRank.DoRank
// (base_general_rank.go:66-109): pre-rank with GeneralRankConfs[scene].RankConf, then the Actions (sort / filter by
// registered name, action.go:61-83) — the second caller of the same rank kernels, over the whole recall set.
std::vector<module::ItemPtr> Rank(module::User* user, std::vector<module::ItemPtr> items, context::RecommendContext* ctx);
}

// ------------------------------------------------------------------------------------------------ sort
namespace sort {
struct SortData {  // sort/sort.go:27-31
  std::vector<module::ItemPtr> Data;
  context::RecommendContext* Context = nullptr;
  module::User* User = nullptr;
  std::string PipelineName;
};
struct ISort {  // :33-35
  virtual ~ISort() = default;
  virtual Error Sort(SortData* sortData) = 0;
};
void RegisterSort(const std::string& name, std::shared_ptr<ISort> s);  // first registration wins (:143-150)
Error GetSort(const std::string& name, std::shared_ptr<ISort>* out);
void Load(const recconf::RecommendConfig& c);                            // SortNames -> strategies (:127-137)
void Sort(SortData* sortData, const std::string& tag);                  // SortService.Sort (:65-125)

class ItemRankScoreSort : public ISort {  // sort/item_rank_score.go: sort.Sort(sort.Reverse(ItemScoreSlice))
 public:
  Error Sort(SortData* d) override;
};
class AlgoScoreSort : public ISort {  // sort/algo_score_sort.go
 public:
  explicit AlgoScoreSort(const recconf::SortConfig& c);
  Error Sort(SortData* d) override;
 private:
  std::string sortByField_;
  double switchThreshold_;
};
// The head of doSort (sort/dpp_sort.go:271-300, sort/ssd_sort.go:298-331): the list the reference holds at the moment
// it loads the embeddings — optionally sorted by score (Go sort order) and cut to max(size, CandidateCount), then to
// the MinScorePercent prefix.  It is also what doSort RETURNS on every error path ("return items").
std::vector<module::ItemPtr> DoSortHead(std::vector<module::ItemPtr> items, int size, int candidateCnt,
                                        double minScorePercent, bool alwaysSort);
// loadEmbeddingCache's guard (dpp_sort.go:246-249, ssd_sort.go:270-273): more than `threshold` of the items have no
// embedding -> error -> doSort returns the head list unchanged
bool EmbeddingMissAboveThreshold(size_t missing, size_t total, double threshold);
// DPPSort.Sort (sort/dpp_sort.go:108-167) with the kernel-matrix + greedy part on the GPU (prg_dpp)
class GpuDPPSort : public ISort {
 public:
  GpuDPPSort(const recconf::DPPSortConfig& c, std::shared_ptr<GpuCatalog> cat);
  Error Sort(SortData* d) override;
 private:
  recconf::DPPSortConfig conf_;
  std::shared_ptr<GpuCatalog> cat_;
};
// sort/dpp_sort.go:52-58: hooks return a []float64 per item; DPPConf.EmbeddingHookNames selects them (:362-370)
using EmbeddingHookFunc = std::function<std::vector<double>(context::RecommendContext*, const module::ItemPtr&)>;
void RegisterEmbeddingHook(const std::string& name, EmbeddingHookFunc fn);
// SSDSort.Sort (sort/ssd_sort.go:108-190) with doSort + SSDWithSlidingWindow on the GPU (prg_ssd)
class GpuSSDSort : public ISort {
 public:
  GpuSSDSort(const recconf::SSDSortConfig& c, std::shared_ptr<GpuCatalog> cat);
  Error Sort(SortData* d) override;
 private:
  recconf::SSDSortConfig conf_;
  std::shared_ptr<GpuCatalog> cat_;
};
}  // namespace sort

namespace filter {
struct FilterData { std::vector<module::ItemPtr> Data; std::string Uid; context::RecommendContext* Context = nullptr; };
struct IFilter {  // filter/filter.go:32-34
  virtual ~IFilter() = default;
  virtual Error Filter(FilterData* d) = 0;
};
void RegisterFilter(const std::string& name, std::shared_ptr<IFilter> f);
Error GetFilter(const std::string& name, std::shared_ptr<IFilter>* out);
void Load(const recconf::RecommendConfig& c);            // UniqueFilter + FilterConfs (AdjustCountFilter)
void UniqueFilter(std::vector<module::ItemPtr>* items);  // filter/unique_filter.go:26-49
}

// ------------------------------------------------------------------------------------------------ table ingest formats
namespace ingest {
// "{v1,v2,...}" embedding text of the Hologres tables DPP reads (sort/dpp_sort.go:224-233): Trim "{}", Split by
// the separator, ParseFloat each element (an unparsable element stays 0, as upstream logs and continues).
std::vector<double> ParseEmbeddingText(const std::string& text, const std::string& sep);
// recall result cache "id:name:score,id:name:score" (service/recall/vector_recall.go:35-58 read, :103-110 write)
std::string FormatFloatV(double v);   // fmt's %v of a float64 (strconv 'g', shortest)
std::string FormatRecallCache(const std::vector<module::ItemPtr>& items, const std::string& modelName);
std::vector<module::ItemPtr> ParseRecallCache(const std::string& s, const std::string& modelName, const std::string& itemType);

// One element of the user-vector text "i:v i:v" (service/recall/vector_recall.go:72-82): strconv.ParseFloat(v, 32) with
// the error ignored — correctly rounded to float32 in one step, 0 when v does not parse as a whole.
float ParseFloat32(const std::string& s);

// utils.ToString (utils/type.go:120-140) for the value kinds a fetched column can hold: integers in decimal, floats as
// strconv.FormatFloat(v, 'f', -1, 64) (shortest digits that round-trip, no exponent), strings unchanged.
std::string ToString(const module::Value& v);

// Item-feature column sets -> the id-encoded field matrix of prg_set_item_fields.  A FeatureDao writes one property per
// non-NULL feature column of a fetched row into item.Properties (module/feature_hologres_dao.go:644-675,
// sqlutil.ParseColumnValues); the reference then ships those VALUES to the remote model, which owns the
// value -> embedding-row mapping.  With the tables in HBM that mapping is explicit, one rule per field:
//   Vocab  the listed values, in order, are rows 0, 1, ... of the field's table (looked up by utils.ToString of the value)
//   IsId   the column already holds the table row (a non-negative integer)
// A NULL / missing column, a value outside the vocabulary and an invalid id encode as kAbsent (0xFFFFFFFF): the gather
// treats an id beyond the table as "no contribution".
struct FieldSpec {
  std::string Column;
  bool IsId = false;
  std::vector<std::string> Vocab;
};
class FieldEncoder {
 public:
  static constexpr uint32_t kAbsent = 0xFFFFFFFFu;
  explicit FieldEncoder(std::vector<FieldSpec> specs);
  size_t size() const { return specs_.size(); }
  void Encode(const module::Features& properties, uint32_t* out) const;   // out[size()]
 private:
  std::vector<FieldSpec> specs_;
  std::vector<std::unordered_map<std::string, uint32_t>> rows_;
};
}

// ------------------------------------------------------------------------------------------------ service
namespace service {
// UserRecommendService.Recommend (service/user_recommend.go:46-183): recall -> filter -> rank -> sort -> items[:size]
std::vector<module::ItemPtr> Recommend(module::User* user, context::RecommendContext* ctx);
}

// Process-wide registries are reset between test cases.
void ResetRegistries();

}  // namespace pairec
