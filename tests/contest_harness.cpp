// TEST / BENCH INFRASTRUCTURE -- the role of the reference's harness (tests/read_sql.cpp:1224-1249):
// builds Plans whose inputs are individually `new`-ed pages through the reference's own
// ColumnInserter (include/plan.h:151-228), calls `Contest::execute` of this repo's drop-in
// (radix-join_b200/libcontest_b200.so) and, for parity, of the UNMODIFIED reference
// (oracle/_ref/libref_oracle.so), and compares the results as sorted multisets after decoding them with
// the reference's Table::from_columnar -- what tests/read_sql.cpp:1159-1222 does.
//
// Compiled against the reference's headers (radix-join_b200/csrc/Makefile, target `contest`), so Plan /
// ColumnarTable / Column / Page are the reference's own types; both libraries are dlopen'ed with
// RTLD_LOCAL because each defines namespace Contest.
//
//   contest_harness parity <c1|c2> [--div D] [--threads T]
//   contest_harness bench  <c1|c2> [--div D] [--steps K] [--warmup W] [--impl ours|reference] [--no-check]
//
// Workloads: c1 = INT32 join, 1 M x 10 M unique foreign keys; c2 = INT32 join 64 Mi x
// 512 Mi, Zipf(0.75) probe keys, INT64 + FP64 payloads with 1 % NULLs; --div divides the row counts.
// One JSON line on stdout; exit code 0 iff everything asked for held.
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include <plan.h>
#include <table.h>

namespace {

using Clock = std::chrono::steady_clock;
double ms_since(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }

struct Impl {
    void* handle = nullptr;
    void* (*build_context)() = nullptr;
    void (*destroy_context)(void*) = nullptr;
    ColumnarTable (*execute)(const Plan&, void*) = nullptr;
    std::string path;
};

Impl load_impl(const std::string& path) {
    Impl im;
    im.path = path;
    im.handle = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!im.handle) {
        std::fprintf(stderr, "dlopen(%s): %s\n", path.c_str(), dlerror());
        std::exit(2);
    }
    // Itanium-mangled names of include/plan.h:337-344
    im.build_context = reinterpret_cast<void* (*)()>(dlsym(im.handle, "_ZN7Contest13build_contextEv"));
    im.destroy_context = reinterpret_cast<void (*)(void*)>(dlsym(im.handle, "_ZN7Contest15destroy_contextEPv"));
    im.execute = reinterpret_cast<ColumnarTable (*)(const Plan&, void*)>(dlsym(im.handle, "_ZN7Contest7executeERK4PlanPv"));
    if (!im.build_context || !im.destroy_context || !im.execute) {
        std::fprintf(stderr, "%s does not export Contest::build_context/destroy_context/execute\n", path.c_str());
        std::exit(2);
    }
    return im;
}

std::string exe_dir() {
    char    buf[4096];
    ssize_t n = readlink("/proc/self/exe", buf, sizeof buf - 1);
    if (n <= 0) return ".";
    buf[n] = 0;
    std::string s(buf);
    return s.substr(0, s.rfind('/'));
}

template <class F>
void parallel_ranges(int threads, uint64_t n, F fn) {
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) {
        const uint64_t b = n * t / threads, e = n * (t + 1) / threads;
        th.emplace_back([=] { fn(b, e, t); });
    }
    for (auto& x: th) x.join();
}

inline uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// YCSB / Gray et al. Zipfian sampler: rank in [0, n), P(rank = i) ~ 1 / (i + 1)^theta
struct Zipf {
    uint64_t n;
    double   theta, zetan, zeta2, alpha, eta;
    Zipf(uint64_t n_, double theta_, int threads): n(n_), theta(theta_) {
        std::vector<double> part(threads, 0.0);
        parallel_ranges(threads, n, [&](uint64_t b, uint64_t e, int t) {
            double acc = 0;
            for (uint64_t i = b; i < e; ++i) acc += std::pow(static_cast<double>(i + 1), -theta);
            part[t] = acc;
        });
        zetan = 0;
        for (double p: part) zetan += p;
        zeta2 = 1.0 + std::pow(0.5, theta);
        alpha = 1.0 / (1.0 - theta);
        eta = (1.0 - std::pow(2.0 / static_cast<double>(n), 1.0 - theta)) / (1.0 - zeta2 / zetan);
    }
    uint64_t rank(double u) const {
        const double uz = u * zetan;
        if (uz < 1.0) return 0;
        if (uz < zeta2) return 1;
        const double r = std::floor(static_cast<double>(n) * std::pow(eta * u - eta + 1.0, alpha));
        if (r < 0) return 0;
        const uint64_t k = static_cast<uint64_t>(r);
        return k >= n ? n - 1 : k;
    }
};

// ---- the synthetic tables -----------------------------------------------------------------------------
// Values are pure functions of the row number (and the permutation), so the generator can also state the
// expected result without joining.
struct Workload {
    bool     payload = false; // c2
    uint64_t n_build = 0, n_probe = 0;
    std::vector<int32_t> perm;          // build keys by build row
    std::vector<int32_t> probe_keys;    // by probe row
    static bool a_valid(uint64_t build_row) { return (splitmix64(build_row * 2 + 1043) >> 11 & 0xFFFFF) >= (1u << 20) / 100; }
    static bool b_valid(uint64_t probe_row) { return (splitmix64(probe_row * 2 + 2044) >> 11 & 0xFFFFF) >= (1u << 20) / 100; }
    static int64_t a_of(int32_t k) { return static_cast<int64_t>(splitmix64(static_cast<uint64_t>(static_cast<int64_t>(k)))); }
    static double  b_of(uint64_t probe_row) {
        uint64_t bits = splitmix64(probe_row);
        if (((bits >> 52) & 0x7FF) == 0x7FF) bits &= ~(1ull << 62); // finite
        double d;
        std::memcpy(&d, &bits, 8);
        return d;
    }
};

// one column from per-thread row ranges: every range is inserted by its own ColumnInserter into its own
// Column, then the page lists are appended in range order (a range's last page may be partly filled)
template <class T, class F>
Column make_column(DataType type, uint64_t n_rows, int threads, F cell /* (row, T* value) -> valid */) {
    std::vector<Column> parts;
    for (int t = 0; t < threads; ++t) parts.emplace_back(type);
    parallel_ranges(threads, n_rows, [&](uint64_t b, uint64_t e, int t) {
        ColumnInserter<T> ins(parts[t]);
        for (uint64_t i = b; i < e; ++i) {
            T v;
            if (cell(i, &v)) ins.insert(v); else ins.insert_null();
        }
        ins.finalize();
    });
    Column out(type);
    for (auto& p: parts) {
        out.pages.insert(out.pages.end(), p.pages.begin(), p.pages.end());
        p.pages.clear();
    }
    return out;
}

Workload make_workload(const std::string& name, uint64_t div, int threads) {
    Workload w;
    w.payload = name == "c2";
    if (name == "c1") {
        w.n_build = 1000000 / div;
        w.n_probe = 10000000 / div;
    } else if (name == "c2") {
        w.n_build = (uint64_t(1) << 26) / div;
        w.n_probe = (uint64_t(1) << 29) / div;
    } else {
        std::fprintf(stderr, "unknown workload %s\n", name.c_str());
        std::exit(2);
    }
    w.perm.resize(w.n_build);
    for (uint64_t i = 0; i < w.n_build; ++i) w.perm[i] = static_cast<int32_t>(i);
    std::mt19937_64 rng(w.payload ? 43 : 42);
    std::shuffle(w.perm.begin(), w.perm.end(), rng);
    w.probe_keys.resize(w.n_probe);
    if (!w.payload) {
        // uniform draws from the same RNG stream
        std::uniform_int_distribution<uint64_t> pick(0, w.n_build - 1);
        for (uint64_t i = 0; i < w.n_probe; ++i) w.probe_keys[i] = static_cast<int32_t>(pick(rng));
    } else {
        Zipf z(w.n_build, 0.75, threads);
        parallel_ranges(threads, w.n_probe, [&](uint64_t b, uint64_t e, int) {
            for (uint64_t i = b; i < e; ++i) {
                const double u = static_cast<double>(splitmix64(i ^ 0x44ull << 56) >> 11) * (1.0 / 9007199254740992.0);
                w.probe_keys[i] = w.perm[z.rank(u)];
            }
        });
    }
    return w;
}

Plan make_plan(const Workload& w, int threads) {
    Plan plan;
    ColumnarTable R, S;
    R.num_rows = w.n_build;
    S.num_rows = w.n_probe;
    R.columns.push_back(make_column<int32_t>(DataType::INT32, w.n_build, threads, [&](uint64_t i, int32_t* v) { *v = w.perm[i]; return true; }));
    S.columns.push_back(make_column<int32_t>(DataType::INT32, w.n_probe, threads, [&](uint64_t i, int32_t* v) { *v = w.probe_keys[i]; return true; }));
    if (w.payload) {
        R.columns.push_back(make_column<int64_t>(DataType::INT64, w.n_build, threads, [&](uint64_t i, int64_t* v) {
            *v = Workload::a_of(w.perm[i]);
            return Workload::a_valid(i);
        }));
        S.columns.push_back(make_column<double>(DataType::FP64, w.n_probe, threads, [&](uint64_t i, double* v) {
            *v = Workload::b_of(i);
            return Workload::b_valid(i);
        }));
    }
    plan.new_input(std::move(R));
    plan.new_input(std::move(S));
    if (w.payload) {
        plan.new_scan_node(0, {{0, DataType::INT32}, {1, DataType::INT64}});
        plan.new_scan_node(1, {{0, DataType::INT32}, {1, DataType::FP64}});
        plan.new_join_node(true, 0, 1, 0, 0, {{0, DataType::INT32}, {1, DataType::INT64}, {3, DataType::FP64}});
    } else {
        // the shape of tests/unit_tests.cpp:12-14
        plan.new_scan_node(0, {{0, DataType::INT32}});
        plan.new_scan_node(1, {{0, DataType::INT32}});
        plan.new_join_node(true, 0, 1, 0, 0, {{0, DataType::INT32}, {1, DataType::INT32}});
    }
    plan.root = 2;
    return plan;
}

// ---- exact multiset comparison (tests/read_sql.cpp:1206-1221: sort both, compare) ---------------------
// rows of <= 3 fixed-width cells become (null mask, bit patterns) so that 10 M-row results sort in seconds
using RowKey = std::array<uint64_t, 4>;

bool rows_of(const ColumnarTable& t, std::vector<RowKey>* out, std::string* why) {
    if (t.columns.size() > 3) {
        *why = "more than 3 columns";
        return false;
    }
    Table decoded = Table::from_columnar(t); // the reference's decoder; throws on malformed pages
    const auto& rows = decoded.table();
    if (rows.size() != t.num_rows) {
        *why = "decoded row count differs from num_rows";
        return false;
    }
    out->resize(rows.size());
    for (size_t i = 0; i < rows.size(); ++i) {
        RowKey k{0, 0, 0, 0};
        for (size_t c = 0; c < rows[i].size(); ++c) {
            const Data& d = rows[i][c];
            uint64_t bits = 0;
            if (std::holds_alternative<std::monostate>(d)) {
                k[0] |= 1ull << c;
            } else if (auto* p32 = std::get_if<int32_t>(&d)) {
                bits = static_cast<uint64_t>(static_cast<int64_t>(*p32));
                k[0] |= 0x10ull << (4 * c);
            } else if (auto* p64 = std::get_if<int64_t>(&d)) {
                bits = static_cast<uint64_t>(*p64);
                k[0] |= 0x20ull << (4 * c);
            } else if (auto* pd = std::get_if<double>(&d)) {
                std::memcpy(&bits, pd, 8);
                k[0] |= 0x30ull << (4 * c);
            } else {
                *why = "VARCHAR cell in a fixed-width workload";
                return false;
            }
            k[c + 1] = bits;
        }
        (*out)[i] = k;
    }
    std::sort(out->begin(), out->end());
    return true;
}

// ---- multiset checksum straight from result pages (full-size runs: no row materialisation) ------------
struct Checksum {
    uint64_t s1 = 0, s2 = 0, rows = 0;
    void add(uint64_t h) {
        s1 += h;
        s2 += splitmix64(h);
        ++rows;
    }
    void merge(const Checksum& o) {
        s1 += o.s1;
        s2 += o.s2;
        rows += o.rows;
    }
    bool operator==(const Checksum& o) const { return s1 == o.s1 && s2 == o.s2 && rows == o.rows; }
};
constexpr uint64_t kNullSalt = 0x6A09E667F3BCC909ull;
inline uint64_t cell_hash(int c, uint64_t bits, bool valid) {
    return valid ? splitmix64(bits ^ (0x9E3779B97F4A7C15ull * static_cast<uint64_t>(c + 1))) : kNullSalt + static_cast<uint64_t>(c);
}
inline uint64_t row_hash(const uint64_t* bits, const bool* valid, int n) {
    uint64_t h = cell_hash(0, bits[0], valid[0]);
    for (int c = 1; c < n; ++c) h = splitmix64(h + cell_hash(c, bits[c], valid[c]));
    return h;
}

// decode rows [r0, r1) of a fixed-width column (page layout: include/plan.h:151-228, src/build_table.cpp:322-380)
void decode_rows(const Column& col, const std::vector<uint64_t>& prefix, uint64_t r0, uint64_t r1, uint64_t* bits, uint8_t* valid) {
    const size_t w = col.type == DataType::INT32 ? 4 : 8;
    size_t p = std::upper_bound(prefix.begin(), prefix.end(), r0) - prefix.begin() - 1;
    uint64_t out = 0;
    while (r0 + out < r1) {
        const uint8_t* pg = reinterpret_cast<const uint8_t*>(col.pages[p]->data);
        uint16_t n_r;
        std::memcpy(&n_r, pg, 2);
        const uint8_t* bm = pg + PAGE_SIZE - (n_r + 7) / 8;
        const uint8_t* vals = pg + (w == 4 ? 4 : 8);
        uint64_t vi = 0;
        const uint64_t first = prefix[p];
        for (uint32_t i = 0; i < n_r && first + i < r1; ++i) {
            const bool ok = (bm[i >> 3] >> (i & 7)) & 1;
            if (first + i >= r0) {
                uint64_t b = 0;
                if (ok) {
                    if (w == 4) {
                        int32_t v;
                        std::memcpy(&v, vals + vi * 4, 4);
                        b = static_cast<uint64_t>(static_cast<int64_t>(v));
                    } else {
                        std::memcpy(&b, vals + vi * 8, 8);
                    }
                }
                bits[first + i - r0] = b;
                valid[first + i - r0] = ok;
                ++out;
            }
            vi += ok;
        }
        ++p;
    }
}

bool result_checksum(const ColumnarTable& t, int threads, Checksum* sum, std::string* why) {
    const int nc = static_cast<int>(t.columns.size());
    std::vector<std::vector<uint64_t>> prefix(nc);
    for (int c = 0; c < nc; ++c) {
        if (t.columns[c].type == DataType::VARCHAR) {
            *why = "VARCHAR column";
            return false;
        }
        auto& pre = prefix[c];
        pre.assign(t.columns[c].pages.size() + 1, 0);
        for (size_t p = 0; p < t.columns[c].pages.size(); ++p) {
            uint16_t n_r;
            std::memcpy(&n_r, t.columns[c].pages[p]->data, 2);
            pre[p + 1] = pre[p] + n_r;
        }
        if (pre.back() != t.num_rows) {
            *why = "column " + std::to_string(c) + " holds " + std::to_string(pre.back()) + " rows, num_rows = " + std::to_string(t.num_rows);
            return false;
        }
    }
    const uint64_t block = 1 << 16;
    const uint64_t n_blocks = (t.num_rows + block - 1) / block;
    std::vector<Checksum> part(threads);
    std::atomic<uint64_t> next{0};
    std::vector<std::thread> th;
    for (int w = 0; w < threads; ++w) {
        th.emplace_back([&, w] {
            std::vector<uint64_t> bits(static_cast<size_t>(nc) * block);
            std::vector<uint8_t>  valid(static_cast<size_t>(nc) * block);
            for (;;) {
                const uint64_t b = next.fetch_add(1);
                if (b >= n_blocks) break;
                const uint64_t r0 = b * block, r1 = std::min(t.num_rows, r0 + block);
                for (int c = 0; c < nc; ++c)
                    decode_rows(t.columns[c], prefix[c], r0, r1, bits.data() + c * block, valid.data() + c * block);
                for (uint64_t i = 0; i < r1 - r0; ++i) {
                    uint64_t rb[3];
                    bool     rv[3];
                    for (int c = 0; c < nc; ++c) {
                        rb[c] = bits[c * block + i];
                        rv[c] = valid[c * block + i] != 0;
                    }
                    part[w].add(row_hash(rb, rv, nc));
                }
            }
        });
    }
    for (auto& x: th) x.join();
    for (auto& p: part) sum->merge(p);
    return true;
}

std::vector<uint64_t> row_prefix_of(const Column& col) {
    std::vector<uint64_t> pre(col.pages.size() + 1, 0);
    for (size_t p = 0; p < col.pages.size(); ++p) {
        uint16_t n_r;
        std::memcpy(&n_r, col.pages[p]->data, 2);
        pre[p + 1] = pre[p] + n_r;
    }
    return pre;
}

// The join result stated WITHOUT joining: every probe key exists exactly once on the build side (a
// permutation), so the build row of a key is a table lookup.  Payload values are read back from the INPUT
// PAGES, not recomputed: the reference's ColumnInserter<8-byte T>::insert tests for 4 free bytes
// (include/plan.h:216), so on pages that hold NULLs the last value can lose its top bytes to the bitmap --
// whatever the pages say is the input both implementations see.
Checksum expected_checksum(const Workload& w, const Plan& plan, int threads) {
    std::vector<uint32_t> row_of_key(w.n_build);
    for (uint64_t i = 0; i < w.n_build; ++i) row_of_key[static_cast<uint32_t>(w.perm[i])] = static_cast<uint32_t>(i);
    std::vector<uint64_t> a_bits;
    std::vector<uint8_t>  a_ok;
    std::vector<uint64_t> pre_a, pre_b;
    const uint64_t block = 1 << 16;
    if (w.payload) {
        const Column& ca = plan.inputs[0].columns[1];
        pre_a = row_prefix_of(ca);
        pre_b = row_prefix_of(plan.inputs[1].columns[1]);
        a_bits.resize(w.n_build);
        a_ok.resize(w.n_build);
        const uint64_t nb = (w.n_build + block - 1) / block;
        parallel_ranges(threads, nb, [&](uint64_t b0, uint64_t b1, int) {
            for (uint64_t b = b0; b < b1; ++b) {
                const uint64_t r0 = b * block, r1 = std::min(w.n_build, r0 + block);
                decode_rows(ca, pre_a, r0, r1, a_bits.data() + r0, a_ok.data() + r0);
            }
        });
    }
    std::vector<Checksum> part(threads);
    const uint64_t n_blocks = (w.n_probe + block - 1) / block;
    parallel_ranges(threads, n_blocks, [&](uint64_t b0, uint64_t b1, int t) {
        std::vector<uint64_t> b_bits(block);
        std::vector<uint8_t>  b_ok(block);
        for (uint64_t b = b0; b < b1; ++b) {
            const uint64_t r0 = b * block, r1 = std::min(w.n_probe, r0 + block);
            if (w.payload) decode_rows(plan.inputs[1].columns[1], pre_b, r0, r1, b_bits.data(), b_ok.data());
            for (uint64_t i = r0; i < r1; ++i) {
                const int32_t k = w.probe_keys[i];
                uint64_t bits[3];
                bool     valid[3];
                bits[0] = static_cast<uint64_t>(static_cast<int64_t>(k));
                valid[0] = true;
                if (w.payload) {
                    const uint32_t br = row_of_key[static_cast<uint32_t>(k)];
                    bits[1] = a_bits[br];
                    valid[1] = a_ok[br] != 0;
                    bits[2] = b_bits[i - r0];
                    valid[2] = b_ok[i - r0] != 0;
                    part[t].add(row_hash(bits, valid, 3));
                } else {
                    bits[1] = bits[0];
                    valid[1] = true;
                    part[t].add(row_hash(bits, valid, 2));
                }
            }
        }
    });
    Checksum s;
    for (auto& p: part) s.merge(p);
    return s;
}

uint64_t total_pages(const ColumnarTable& t) {
    uint64_t n = 0;
    for (auto& c: t.columns) n += c.pages.size();
    return n;
}

} // namespace

int main(int argc, char** argv) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: contest_harness parity|bench c1|c2 [--div D] [--steps K] [--warmup W] [--impl ours|reference] [--threads T] [--no-check]\n");
        return 2;
    }
    const std::string mode = argv[1], workload = argv[2];
    uint64_t    div = 1;
    int         steps = 3, warmup = 1, threads = static_cast<int>(std::thread::hardware_concurrency());
    std::string impl = "ours";
    bool        check = true;
    for (int i = 3; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&] { return i + 1 < argc ? std::string(argv[++i]) : std::string("0"); };
        if (a == "--div") div = std::stoull(next());
        else if (a == "--steps") steps = std::stoi(next());
        else if (a == "--warmup") warmup = std::stoi(next());
        else if (a == "--impl") impl = next();
        else if (a == "--threads") threads = std::stoi(next());
        else if (a == "--no-check") check = false;
    }
    if (threads < 1) threads = 1;
    if (div < 1) div = 1;
    const std::string here = exe_dir(); // oracle/_ref
    const std::string ours_path = here + "/../../radix-join_b200/libcontest_b200.so";
    const std::string ref_path = here + "/libref_oracle.so";

    auto t0 = Clock::now();
    Workload w = make_workload(workload, div, threads);
    Plan     plan = make_plan(w, threads);
    const double gen_ms = ms_since(t0);
    uint64_t in_pages = 0;
    for (auto& t: plan.inputs) in_pages += total_pages(t);

    if (mode == "parity") {
        Impl ours = load_impl(ours_path), ref = load_impl(ref_path);
        void* ctx = ours.build_context();
        t0 = Clock::now();
        ColumnarTable got = ours.execute(plan, ctx);
        const double ours_ms = ms_since(t0);
        // a second call reuses the context's staging buffers and worker pool
        ColumnarTable got2 = ours.execute(plan, ctx);
        ours.destroy_context(ctx);
        void* rctx = ref.build_context();
        t0 = Clock::now();
        ColumnarTable want = ref.execute(plan, rctx);
        const double ref_ms = ms_since(t0);
        ref.destroy_context(rctx);
        std::string why;
        bool ok = got.num_rows == want.num_rows && got.columns.size() == want.columns.size() && got2.num_rows == want.num_rows;
        for (size_t c = 0; ok && c < got.columns.size(); ++c) ok = got.columns[c].type == want.columns[c].type;
        if (!ok) why = "row count / column types differ";
        std::vector<RowKey> a, b, a2;
        if (ok) ok = rows_of(got, &a, &why) && rows_of(want, &b, &why) && rows_of(got2, &a2, &why);
        if (ok && (a != b || a2 != b)) {
            ok = false;
            why = "sorted multisets differ";
        }
        // the generator's own statement of the result must agree as well
        Checksum sum, exp = expected_checksum(w, plan, threads);
        if (ok) {
            ok = result_checksum(got, threads, &sum, &why);
            if (ok && !(sum == exp)) {
                ok = false;
                why = "multiset checksum differs from the generator's expectation";
            }
        }
        std::printf("{\"mode\": \"parity\", \"workload\": \"%s\", \"div\": %llu, \"build_rows\": %llu, \"probe_rows\": %llu, \"input_pages\": %llu, "
                    "\"rows\": %llu, \"ref_rows\": %llu, \"output_pages\": %llu, \"ours_ms\": %.2f, \"reference_ms\": %.2f, \"ok\": %s, \"why\": \"%s\"}\n",
                    workload.c_str(), (unsigned long long)div, (unsigned long long)w.n_build, (unsigned long long)w.n_probe,
                    (unsigned long long)in_pages, (unsigned long long)got.num_rows, (unsigned long long)want.num_rows,
                    (unsigned long long)total_pages(got), ours_ms, ref_ms, ok ? "true" : "false", why.c_str());
        return ok ? 0 : 1;
    }

    if (mode == "bench") {
        Impl im = load_impl(impl == "reference" ? ref_path : ours_path);
        t0 = Clock::now();
        void* ctx = im.build_context();
        const double ctx_ms = ms_since(t0);
        std::vector<double> times;
        uint64_t rows = 0, out_pages = 0;
        bool     ok = true;
        std::string why;
        Checksum exp;
        if (check) exp = expected_checksum(w, plan, threads);
        for (int i = 0; i < warmup + steps; ++i) {
            // timed exactly like the contest harness: steady_clock around Contest::execute (tests/read_sql.cpp:1234-1236)
            t0 = Clock::now();
            ColumnarTable res = im.execute(plan, ctx);
            const double ms = ms_since(t0);
            if (i >= warmup) times.push_back(ms);
            rows = res.num_rows;
            out_pages = total_pages(res);
            if (check && i == warmup + steps - 1) {
                Checksum sum;
                ok = result_checksum(res, threads, &sum, &why);
                if (ok && !(sum == exp)) {
                    ok = false;
                    why = "multiset checksum differs from the generator's expectation";
                }
            }
            // the result is destroyed here, before the next call, as the contest harness does
        }
        im.destroy_context(ctx);
        double sum = 0, best = 1e300;
        for (double t: times) {
            sum += t;
            best = std::min(best, t);
        }
        const double mean = times.empty() ? 0 : sum / times.size();
        std::printf("{\"mode\": \"bench\", \"impl\": \"%s\", \"workload\": \"%s\", \"div\": %llu, \"build_rows\": %llu, \"probe_rows\": %llu, "
                    "\"input_pages\": %llu, \"rows\": %llu, \"output_pages\": %llu, \"steps\": %d, \"warmup\": %d, \"ms_mean\": %.3f, \"ms_best\": %.3f, "
                    "\"build_context_ms\": %.2f, \"generate_ms\": %.1f, \"threads\": %d, \"checked\": %s, \"ok\": %s, \"why\": \"%s\"}\n",
                    impl.c_str(), workload.c_str(), (unsigned long long)div, (unsigned long long)w.n_build, (unsigned long long)w.n_probe,
                    (unsigned long long)in_pages, (unsigned long long)rows, (unsigned long long)out_pages, steps, warmup, mean, best, ctx_ms, gen_ms,
                    threads, check ? "true" : "false", ok ? "true" : "false", why.c_str());
        return ok ? 0 : 1;
    }
    std::fprintf(stderr, "unknown mode %s\n", mode.c_str());
    return 2;
}
