// Stand-in for googletest (v1.17.0 in the reference, third party, cmake FetchContent), on the include path of the
// oracle/_ref/ref_gtests build only: just enough of TEST / EXPECT_* / ASSERT_* to compile the reference's own test files
// unmodified and run them (oracle/ref_gtests_main.cu).  A failed ASSERT_* returns from the test body like gtest's does.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

namespace testing_shim {
struct Test
{
    const char* suite;
    const char* name;
    void (*body)();
};
inline std::vector<Test>& registry()
{
    static std::vector<Test> r;
    return r;
}
inline int& failures()
{
    static int f = 0;
    return f;
}
struct Registrar
{
    Registrar(const char* s, const char* n, void (*b)()) { registry().push_back({s, n, b}); }
};
template <typename T>
std::string show(const T& v)
{
    if constexpr (std::is_arithmetic_v<T> || std::is_enum_v<T>) {
        std::ostringstream os;
        if constexpr (std::is_enum_v<T>)
            os << (long long)v;
        else
            os << +v;
        return os.str();
    } else {
        return "<value>";
    }
}
inline void fail(const char* file, int line, const std::string& what)
{
    ++failures();
    fprintf(stderr, "%s:%d: Failure\n%s\n", file, line, what.c_str());
}
// 4-ULP comparison like EXPECT_FLOAT_EQ
inline bool almost_equal(float a, float b)
{
    if (std::isnan(a) || std::isnan(b)) return false;
    if (a == b) return true;
    int32_t ia, ib;
    memcpy(&ia, &a, 4), memcpy(&ib, &b, 4);
    if (ia < 0) ia = (int32_t)0x80000000 - ia;
    if (ib < 0) ib = (int32_t)0x80000000 - ib;
    return std::llabs((long long)ia - (long long)ib) <= 4;
}
}  // namespace testing_shim

#define TEST(suite, name)                                                                                           \
    static void suite##_##name##_body();                                                                            \
    static testing_shim::Registrar suite##_##name##_registrar(#suite, #name, &suite##_##name##_body);               \
    static void                    suite##_##name##_body()

#define RXM_GT_CHECK(cond, text, on_fail)                                              \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            testing_shim::fail(__FILE__, __LINE__, std::string("Expected: ") + text);  \
            on_fail;                                                                   \
        }                                                                              \
    } while (0)
#define RXM_GT_CMP(a, op, b, on_fail)                                                                                       \
    do {                                                                                                                    \
        auto&& rxm_gt_a = (a);                                                                                              \
        auto&& rxm_gt_b = (b);                                                                                              \
        if (!(rxm_gt_a op rxm_gt_b)) {                                                                                      \
            testing_shim::fail(__FILE__, __LINE__, std::string("Expected: (" #a ") " #op " (" #b "), actual: ") +          \
                                                       testing_shim::show(rxm_gt_a) + " vs " + testing_shim::show(rxm_gt_b)); \
            on_fail;                                                                                                        \
        }                                                                                                                   \
    } while (0)

#define EXPECT_TRUE(c) RXM_GT_CHECK((c), #c " is true", (void)0)
#define EXPECT_FALSE(c) RXM_GT_CHECK(!(c), #c " is false", (void)0)
#define ASSERT_TRUE(c) RXM_GT_CHECK((c), #c " is true", return)
#define ASSERT_FALSE(c) RXM_GT_CHECK(!(c), #c " is false", return)
#define EXPECT_EQ(a, b) RXM_GT_CMP(a, ==, b, (void)0)
#define EXPECT_NE(a, b) RXM_GT_CMP(a, !=, b, (void)0)
#define EXPECT_LT(a, b) RXM_GT_CMP(a, <, b, (void)0)
#define EXPECT_LE(a, b) RXM_GT_CMP(a, <=, b, (void)0)
#define EXPECT_GT(a, b) RXM_GT_CMP(a, >, b, (void)0)
#define EXPECT_GE(a, b) RXM_GT_CMP(a, >=, b, (void)0)
#define ASSERT_EQ(a, b) RXM_GT_CMP(a, ==, b, return)
#define ASSERT_NE(a, b) RXM_GT_CMP(a, !=, b, return)
#define EXPECT_FLOAT_EQ(a, b) RXM_GT_CHECK(testing_shim::almost_equal((float)(a), (float)(b)), #a " ~= " #b, (void)0)
#define ASSERT_FLOAT_EQ(a, b) RXM_GT_CHECK(testing_shim::almost_equal((float)(a), (float)(b)), #a " ~= " #b, return)
#define EXPECT_NEAR(a, b, tol) RXM_GT_CHECK(std::fabs((double)(a) - (double)(b)) <= (double)(tol), #a " near " #b, (void)0)
#define ASSERT_NEAR(a, b, tol) RXM_GT_CHECK(std::fabs((double)(a) - (double)(b)) <= (double)(tol), #a " near " #b, return)
#define EXPECT_STREQ(a, b) RXM_GT_CHECK(std::string(a) == std::string(b), #a " == " #b, (void)0)
