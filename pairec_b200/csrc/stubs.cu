// stubs.cu — entry points not implemented yet (replaced one by one; file goes away when empty).
#include "handle.h"
using namespace prg;
#define NYI(name) return prg::fail(PRG_EUNSUPPORTED, name " not implemented yet")
extern "C" {
int prg_set_item_fields(prg_handle*, const uint32_t*, uint64_t, uint32_t, int) { NYI("prg_set_item_fields"); }
int prg_set_feature_table(prg_handle*, int, const float*, const float*, uint64_t, uint32_t, int) { NYI("prg_set_feature_table"); }
int prg_set_fm_bias(prg_handle*, float) { NYI("prg_set_fm_bias"); }
int prg_set_mlp(prg_handle*, int, const uint32_t*, const uint16_t* const*, const float* const*) { NYI("prg_set_mlp"); }
int prg_set_diversity_matrix(prg_handle*, const void*, uint64_t, uint32_t, int, int) { NYI("prg_set_diversity_matrix"); }
int prg_rank(prg_handle*, int, const uint32_t*, int, int, double*, int) { NYI("prg_rank"); }
int prg_sort_desc_host(const double*, int, int32_t*) { NYI("prg_sort_desc_host"); }
int prg_sort_desc(prg_handle*, const double*, int, int, int32_t*, int) { NYI("prg_sort_desc"); }
int prg_dpp(prg_handle*, const uint32_t*, const double*, int, int, const prg_dpp_params*, int32_t*, int32_t*, int32_t*, int) { NYI("prg_dpp"); }
int prg_recommend(prg_handle*, const float*, int, int, int, const prg_dpp_params*, uint32_t*, double*, int32_t*, int) { NYI("prg_recommend"); }
}
