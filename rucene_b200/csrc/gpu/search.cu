// search.cu — (stub, replaced by the real planner) batch search entry points.
#include "engine.hpp"
using namespace rg;
extern "C" {
int rg_search_batch(rg_engine*, const rg_query*, uint32_t, const rg_clause*, uint32_t, const rg_search_params*, rg_hit*, uint32_t*, uint64_t*) { g_last_error = "not built yet"; return RG_EUNSUPPORTED; }
int rg_batch_prepare(rg_engine*, const rg_query*, uint32_t, const rg_clause*, uint32_t, const rg_search_params*, rg_batch**) { return RG_EUNSUPPORTED; }
int rg_batch_run(rg_engine*, rg_batch*) { return RG_EUNSUPPORTED; }
int rg_batch_fetch(rg_engine*, rg_batch*, rg_hit*, uint32_t*, uint64_t*) { return RG_EUNSUPPORTED; }
void rg_batch_destroy(rg_engine*, rg_batch*) {}
int rg_batch_stats(rg_engine*, rg_batch*, uint64_t*) { return RG_EUNSUPPORTED; }
int rg_batch_leaf_records(rg_engine*, rg_batch*, void**, size_t*) { return RG_EUNSUPPORTED; }
int rg_merge_leaf_records(rg_engine*, const void*, uint32_t, uint32_t, uint32_t, rg_hit*, uint32_t*, uint64_t*) { return RG_EUNSUPPORTED; }
}
