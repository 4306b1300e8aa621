// The reference's examples/example.rs:111-117 against the C++ host mirror (searcher.hpp):
// builds a small synthetic segment with librucene_codec, searches it on the GPU, prints TopDocs
// as "total_hits\n doc score_bits ..." so the pytest driver can compare with the oracle.
#include <cstdio>
#include <cstring>

#include "../../rucene_b200/csrc/host/searcher.hpp"
#include "rucene_codec.h"

int main() {
    using namespace rucene;
    rc_synth_config cfg{0x5EED0001ull, 50000, 500, 1, 2};
    rc_segment* seg = rc_synth_segment(&cfg);
    if (!seg) { std::fprintf(stderr, "synth failed: %s\n", rc_last_error()); return 2; }
    LeafData leaf;
    leaf.doc_file = rc_segment_doc_file(seg, &leaf.doc_len);
    leaf.norms = rc_segment_norms(seg);
    leaf.terms = rc_segment_terms(seg, &leaf.n_terms);
    int64_t st[8];
    rc_segment_stats(seg, st);
    leaf.doc_count = st[0]; leaf.sum_total_term_freq = st[1]; leaf.sum_doc_freq = st[2]; leaf.max_doc = (int32_t)st[3];
    std::unordered_map<std::string, uint32_t> dict;
    for (uint32_t t = 0; t < leaf.n_terms; t++) dict["t" + std::to_string(t)] = t;
    try {
        GpuIndexSearcher searcher({leaf}, "body", dict);
        auto q1 = TermQuery::create(Term::create("body", "t5"), 1.0f);
        auto q2 = BooleanQuery::build({TermQuery::create(Term::create("body", "t3")), TermQuery::create(Term::create("body", "t40"))}, {}, {}, {}, 0);
        auto q3 = BooleanQuery::build({}, {TermQuery::create(Term::create("body", "t1")), TermQuery::create(Term::create("body", "t77")),
                                           TermQuery::create(Term::create("body", "nope"))}, {}, {}, 0);
        // MUST + SHOULD: ReqOptScorer (boolean_query.rs:253-262)
        auto q4 = BooleanQuery::build({TermQuery::create(Term::create("body", "t2"))},
                                      {TermQuery::create(Term::create("body", "t9")), TermQuery::create(Term::create("body", "t30"))}, {}, {}, 0);
        for (const QueryPtr& q : {q1, q2, q3, q4}) {
            TopDocsCollector collector(10);
            searcher.search(*q, collector);
            const TopDocs& top = collector.top_docs();
            std::printf("%llu", (unsigned long long)top.total_hits());
            for (const ScoreDoc& d : top.score_docs()) {
                uint32_t bits; std::memcpy(&bits, &d.score, 4);
                std::printf(" %d:%u", d.doc_id(), bits);
            }
            std::printf("\n");
        }
        {   // MUST + FILTER on the same term: the filter restricts (here: not at all) and scores 0f32
            auto q5 = BooleanQuery::build({q1}, {}, {q1}, {}, 0);
            TopDocsCollector c(10), c1(10);
            searcher.search(*q5, c);
            searcher.search(*q1, c1);
            bool same = c.top_docs().total_hits() == c1.top_docs().total_hits() &&
                        c.top_docs().score_docs().size() == c1.top_docs().score_docs().size();
            for (size_t i = 0; same && i < c.top_docs().score_docs().size(); i++)
                same = c.top_docs().score_docs()[i].doc == c1.top_docs().score_docs()[i].doc &&
                       c.top_docs().score_docs()[i].score == c1.top_docs().score_docs()[i].score;
            std::printf("filter_same_as_must:%d\n", (int)same);
        }
        bool threw = false;
        try {  // a nested BooleanQuery clause is outside the accelerated path
            auto inner = BooleanQuery::build({q1, q1}, {}, {}, {}, 0);
            auto q6 = BooleanQuery::build({inner, q1}, {}, {}, {}, 0);
            TopDocsCollector c(10);
            searcher.search(*q6, c);
        } catch (const UnsupportedQuery&) { threw = true; }
        std::printf("unsupported:%d\n", (int)threw);
    } catch (const Error& e) {
        std::fprintf(stderr, "error %d: %s\n", e.code, e.what());
        return 1;
    }
    rc_segment_destroy(seg);
    return 0;
}
