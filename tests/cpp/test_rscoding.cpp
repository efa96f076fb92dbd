// test_rscoding.cpp -- the reference's unit tests for this path, restated against the C++ host mirror:
//   src/utils/bitmap.rs:312-420   and   src/utils/rscoding.rs:685-877
// Usage: test_rscoding host   (bookkeeping only; runs without a GPU)
//        test_rscoding gpu    (everything, shard bytes computed by the CUDA kernels through the C ABI)
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <string>

#include "../../summerset_b200/host/summerset_host.hpp"

using namespace ssb;
static int failures = 0;
#define CHECK(cond)                                                           \
    do {                                                                      \
        if (!(cond)) { ++failures; std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)
static bool is_err(const std::function<void()> &f) {
    try { f(); } catch (const SummersetError &) { return true; }
    return false;
}
static bool panics(const std::function<void()> &f) {
    try { f(); } catch (const std::logic_error &) { return true; }
    return false;
}
// bincode(TestData("interesting_value")) = varint length 17 + bytes (rscoding.rs:698)
static RSCodeword::Bytes test_data() {
    std::string s = "interesting_value";
    RSCodeword::Bytes b{static_cast<uint8_t>(s.size())};
    b.insert(b.end(), s.begin(), s.end());
    return b;
}

static void bitmap_tests() {
    CHECK(panics([] { Bitmap b(0, true); }));                                   // new_invalid
    Bitmap ref = Bitmap::from(5, {1, 2, 3});                                    // conversions
    CHECK((ref.to_vec() == std::vector<uint8_t>{1, 2, 3}));
    Bitmap m(7, false);                                                         // bitmap_set_get
    m.set(0, true); m.set(1, false); m.set(2, true);
    CHECK(is_err([&] { m.set(7, true); }));
    CHECK(m.get(0) && !m.get(1) && m.get(2) && !m.get(3));
    CHECK(is_err([&] { m.get(7); }));
    Bitmap f(5, false); f.set(1, true); f.flip();                               // bitmap_flip
    CHECK(f == Bitmap::from(5, {0, 2, 3, 4}));
    Bitmap a = Bitmap::from(5, {0, 1, 3}); a.union_with(Bitmap::from(5, {0, 4})); // bitmap_union
    CHECK(a == Bitmap::from(5, {0, 1, 3, 4}));
    CHECK(is_err([&] { a.union_with(Bitmap(6, false)); }));
    Bitmap c(7, false); CHECK(c.count() == 0);                                  // bitmap_count
    c.set(0, true); c.set(2, true); c.set(3, true); CHECK(c.count() == 3);
    Bitmap it(5, true); it.set(2, false);                                       // bitmap_iter
    CHECK((it.to_vec() == std::vector<uint8_t>{0, 1, 3, 4}));
}

static void host_tests() {
    auto data = test_data();
    const size_t data_len = data.size(), L = data_len % 3 == 0 ? data_len / 3 : data_len / 3 + 1;
    // new_from_data (rscoding.rs:697-736)
    CHECK(is_err([&] { RSCodeword::from_data(data, 0, 0); }));
    RSCodeword cw = RSCodeword::from_data(data, 3, 0);
    CHECK(cw.num_data_shards() == 3 && cw.num_parity_shards() == 0 && cw.num_shards() == 3);
    CHECK(cw.avail_data_shards() == 3 && cw.avail_parity_shards() == 0 && cw.avail_shards() == 3);
    CHECK(cw.avail_shards_map() == Bitmap::from(3, {0, 1, 2}));
    CHECK(cw.data_len() == data_len && cw.shard_len() == L);
    cw = RSCodeword::from_data(data, 3, 2);
    CHECK(cw.num_shards() == 5 && cw.avail_shards() == 3 && cw.avail_shards_map() == Bitmap::from(5, {0, 1, 2}));
    CHECK(cw.data_len() == data_len && cw.shard_len() == L);
    // new_from_null (rscoding.rs:738-754)
    CHECK(is_err([] { RSCodeword::from_null(0, 0); }));
    RSCodeword nul = RSCodeword::from_null(3, 2);
    CHECK(nul.num_shards() == 5 && nul.avail_shards() == 0 && nul.avail_shards_map() == Bitmap(5, false));
    CHECK(nul.data_len() == 0 && nul.shard_len() == 0);
    // subset_absorb (rscoding.rs:756-786)
    RSCodeword cwa = RSCodeword::from_data(data, 3, 2);
    CHECK(is_err([&] { cwa.subset_copy(Bitmap::from(6, {0, 5}), false); }));
    RSCodeword cw01 = cwa.subset_copy(Bitmap::from(5, {0, 1}), false);
    CHECK(cw01.avail_data_shards() == 2);
    RSCodeword cw02 = cwa.subset_copy(Bitmap::from(5, {0, 2}), true);
    CHECK(cw02.avail_data_shards() == 2 && cw02.data_copy.has_value());
    RSCodeword cwb = RSCodeword::from_null(3, 2);
    cwb.absorb_other(cw02);
    CHECK(cwb.avail_shards() == 2 && cwb.avail_shards_map() == Bitmap::from(5, {0, 2}));
    cwb.absorb_other(cw01);
    CHECK(cwb.avail_shards() == 3 && cwb.avail_shards_map() == Bitmap::from(5, {0, 1, 2}));
    CHECK(cwb.get_data() == data);
    CHECK(is_err([&] { cwb.absorb_other(RSCodeword::from_data(data, 5, 3)); }));
    // p == 0 paths never touch a coder (rscoding.rs:454-456,498-507,549-557)
    RSCodeword c0 = RSCodeword::from_data(data, 3, 0);
    c0.compute_parity(nullptr);
    CHECK(c0.avail_parity_shards() == 0 && c0.verify_parity(nullptr));
    c0.reconstruct_all(nullptr);
    c0.shards[1].reset();
    CHECK(is_err([&] { c0.reconstruct_all(nullptr); }) && is_err([&] { c0.reconstruct_data(nullptr); }));
}

static void gpu_tests() {
    auto data = test_data();
    ReedSolomon rs32(3, 2), rs53(5, 3);
    // compute_verify (rscoding.rs:788-818)
    RSCodeword cw_null = RSCodeword::from_null(3, 2);
    CHECK(is_err([&] { cw_null.compute_parity(&rs32); }) && is_err([&] { cw_null.verify_parity(&rs32); }));
    RSCodeword cw_part = RSCodeword::from_data(data, 3, 2);
    cw_part.shards[1].reset();
    CHECK(is_err([&] { cw_part.compute_parity(&rs32); }) && is_err([&] { cw_part.verify_parity(&rs32); }));
    RSCodeword cw = RSCodeword::from_data(data, 3, 2);
    cw.compute_parity(&rs32);
    CHECK(cw.avail_parity_shards() == 2);
    CHECK(cw.verify_parity(&rs32));
    const uint8_t want3[6] = {0x2b, 0x6c, 0x7b, 0x71, 0x7e, 0x70}, want4[6] = {0x2f, 0xfb, 0x9c, 0xcc, 0x5d, 0xa8};
    CHECK(std::memcmp(cw.shards[3]->data(), want3, 6) == 0 && std::memcmp(cw.shards[4]->data(), want4, 6) == 0);
    CHECK(is_err([&] { cw.compute_parity(nullptr); }) && is_err([&] { cw.compute_parity(&rs53); }));
    CHECK(is_err([&] { cw.verify_parity(nullptr); }) && is_err([&] { cw.verify_parity(&rs53); }));
    // reconstruction (rscoding.rs:820-862)
    CHECK(is_err([&] { cw_null.reconstruct_all(&rs32); }) && is_err([&] { cw_null.reconstruct_data(&rs32); }));
    CHECK(is_err([&] { cw_part.reconstruct_all(&rs32); }) && is_err([&] { cw_part.reconstruct_data(&rs32); }));
    RSCodeword r = RSCodeword::from_data(data, 3, 2);
    r.reconstruct_all(&rs32);
    CHECK(r.avail_shards() == 5);
    auto golden = r.shards;
    r.shards[1].reset(); r.shards[3].reset();
    r.reconstruct_all(&rs32);
    CHECK(r.avail_shards() == 5 && r.shards == golden);
    r.shards[0].reset(); r.shards[2].reset();
    r.reconstruct_data(&rs32);
    CHECK(r.avail_data_shards() == 3 && r.shards == golden);
    r.shards[0].reset(); r.shards[1].reset(); r.shards[4].reset();
    CHECK(is_err([&] { r.reconstruct_all(&rs32); }) && is_err([&] { r.reconstruct_data(&rs32); }));
    CHECK(!r.shards[0] && !r.shards[1] && !r.shards[4]);
    CHECK(is_err([&] { r.reconstruct_all(nullptr); }) && is_err([&] { r.reconstruct_all(&rs53); }));
    // get_data (rscoding.rs:864-876)
    RSCodeword g = RSCodeword::from_data(data, 3, 2);
    CHECK(g.get_data() == data);
    g.compute_parity(&rs32);
    g.shards[0].reset(); g.data_copy.reset();
    CHECK(is_err([&] { g.get_data(); }));
    g.reconstruct_data(&rs32);
    CHECK(g.get_data() == data);
    // ReedSolomon::new errors
    CHECK(is_err([] { ReedSolomon bad(0, 1); }) && is_err([] { ReedSolomon bad(3, 0); }) && is_err([] { ReedSolomon bad(200, 57); }));
}

int main(int argc, char **argv) {
    const std::string mode = argc > 1 ? argv[1] : "host";
    bitmap_tests();
    host_tests();
    if (mode == "gpu") gpu_tests();
    else {
        // without a GPU the coder must refuse loudly, never fall back
        bool refused = false;
        try { ReedSolomon rs(3, 2); } catch (const SummersetError &e) { refused = (e.code == SS_ERR_NO_DEVICE); }
        if (std::getenv("SS_EXPECT_NO_GPU")) CHECK(refused);
    }
    std::printf("%s: %d failure(s)\n", mode.c_str(), failures);
    return failures == 0 ? 0 : 1;
}
