// Exercises include/aha_b200.hpp (the C++ twin of aha's InferenceModel seam) without a GPU:
//   * generate_generic semantics (/root/reference/src/models/common/generate.rs:115-159) on a scripted model,
//   * the host-only M-RoPE index through the wrapper,
//   * error propagation of B200Model when no sm_100 device is present (argv[1] == "nogpu").
#include <cstdio>
#include <cstring>
#include <string>

#include "aha_b200.hpp"

#define REQUIRE(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

// A model whose next token is scripted; records how it was driven.
struct Scripted : aha::InferenceModel {
    std::vector<uint32_t> script, eos;
    size_t V = 16, calls = 0, clears = 0;
    std::vector<size_t> offsets, lens;
    std::vector<float> logits_for(uint32_t tok) const {
        std::vector<float> l(V, -1.0f);
        l[tok] = 2.0f;
        l[(tok + 1) % V] = 1.5f;   // runner-up, so a repeat penalty can flip the choice
        return l;
    }
    std::vector<float> forward_initial(const std::vector<uint32_t>& ids, size_t off, const aha::MultiModalData*) override {
        offsets.push_back(off); lens.push_back(ids.size());
        return logits_for(script.at(calls++));
    }
    std::vector<float> forward_step(const std::vector<uint32_t>& ids, size_t off) override {
        offsets.push_back(off); lens.push_back(ids.size());
        return logits_for(script.at(calls++));
    }
    void clear_cache() override { ++clears; }
    std::vector<uint32_t> stop_token_ids() const override { return eos; }
};

int main(int argc, char** argv) {
    using namespace aha;
    {   // EOS handling: the first token is never checked, a later EOS is pushed and ends the loop
        Scripted m; m.script = {7, 3, 7, 5, 5, 5}; m.eos = {7};
        Usage u;
        auto out = generate_generic(m, {1, 2, 3, 4}, nullptr, GenerationContext(std::nullopt, std::nullopt, std::nullopt, 4, 10), &u);
        REQUIRE((out == std::vector<uint32_t>{7, 3, 7}));
        REQUIRE((m.offsets == std::vector<size_t>{0, 4, 5}));       // seqlen_offset: 0, then S, S+1, ...
        REQUIRE((m.lens == std::vector<size_t>{4, 1, 1}));
        REQUIRE(m.clears == 1 && u.prompt_tokens == 4 && u.completion_tokens == 3);
    }
    {   // sample_len caps the number of generated tokens (prefill token included)
        Scripted m; m.script = {1, 2, 3, 4, 5, 6, 7, 8};
        auto out = generate_generic(m, {9}, nullptr, GenerationContext(0.0f, std::nullopt, std::nullopt, 1, 5));
        REQUIRE((out == std::vector<uint32_t>{1, 2, 3, 4, 5}));
        REQUIRE(m.calls == 5);
    }
    {   // repeat penalty (sample.rs:40-60): 2.0 / 1.5 < 1.5, so a repeated top-1 loses to the runner-up
        Scripted m; m.script = {4, 4, 4};
        auto out = generate_generic(m, {9}, nullptr, GenerationContext(std::nullopt, 1.5f, 64, 1, 3));
        REQUIRE((out == std::vector<uint32_t>{4, 5, 4}));           // step 2: 4 penalised -> 5; step 3: 4 (1.33) vs 5 (1.0)
        Scripted k; k.script = {4, 4, 4};
        auto plain = generate_generic(k, {9}, nullptr, GenerationContext(std::nullopt, 1.0f, 64, 1, 3));
        REQUIRE((plain == std::vector<uint32_t>{4, 4, 4}));
    }
    {   // non-greedy requests are rejected like the library does
        bool threw = false;
        try { GenerationContext c(0.7f, std::nullopt, std::nullopt, 1, 4); (void)c; } catch (const Error&) { threw = true; }
        REQUIRE(threw);
    }
    {   // M-RoPE known answer: one 1088x1920 image (grid 1x68x120, merge 2) + 512 text ids -> rope_delta = -1980
        const uint32_t VS = 900, IMG = 901, VE = 902;
        std::vector<uint32_t> ids{VS};
        ids.insert(ids.end(), 2040, IMG);
        ids.push_back(VE);
        for (uint32_t i = 0; i < 512; ++i) ids.push_back(10 + i % 100);
        auto [pos, delta] = rope_index(ids, {1, 68, 120}, 2, IMG, VS);
        REQUIRE(delta == -1980);
        REQUIRE(pos[0] == 0 && pos[1] == 1 && pos[ids.size() + 1] == 1 && pos[2 * ids.size() + 2] == 2);   // image block starts at 1: t=1, h=1.., w=1..
        REQUIRE(pos[ids.size() - 1] == (int32_t)ids.size() - 1 - 1980);
        bool threw = false;
        try { rope_index(ids, {}, 2, IMG, VS); } catch (const Error&) { threw = true; }   // text-only call on an image prompt is still valid (arange)
        REQUIRE(!threw);
        try { rope_index(ids, {1, 68, 120, 1, 4, 4}, 2, IMG, VS); } catch (const Error&) { threw = true; }
        REQUIRE(!threw);   // an unused extra grid row is not an error in the reference either
        std::vector<uint32_t> bad = ids; bad.push_back(VS);
        try { rope_index(bad, {1, 68, 120}, 2, IMG, VS); } catch (const Error& e) { threw = std::strstr(e.what(), "vision_start") != nullptr; }
        REQUIRE(threw);
    }
    {   // the prefix-cache rule (host only): multi-turn reuse, the last prompt token always runs, placeholders must lie inside the prefix
        const uint32_t IMG = 9;
        REQUIRE(prefix_match({1, 2, 3, 4}, {1, 2, 3, 4, 5, 6}) == 4);
        REQUIRE(prefix_match({1, 2, 3, 4}, {1, 2, 3, 4}) == 3);
        REQUIRE(prefix_match({5, IMG, IMG, 6, 7}, {5, IMG, IMG, 6, 8}, {IMG}) == 4);
        REQUIRE(prefix_match({5, IMG, IMG, 6, 7}, {5, IMG, IMG, 6, 8}, {IMG}, false) == 0);
        REQUIRE(prefix_match({5, 6, IMG, IMG}, {5, 7, IMG, IMG}, {IMG}) == 0);
    }
    if (argc > 1 && std::string(argv[1]) == "nogpu") {   // no CPU fallback: the constructor throws with the library's message
        bool threw = false;
        try {
            B200Model m("qwen3", "{}", {});
        } catch (const Error& e) {
            threw = true;
            std::printf("create error: %s\n", e.what());
        }
        REQUIRE(threw);
    }
    std::printf("host mirror OK\n");
    return 0;
}
