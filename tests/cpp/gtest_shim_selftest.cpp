#include <atomic>
#include <cstdint>
#include "gtest/gtest.h"
TEST(Shim, Passes)
{
    std::atomic_uint32_t n{3};
    EXPECT_EQ(n, 3u);
    EXPECT_FLOAT_EQ(std::sqrt(2.0f * 2.0f * 386), std::sqrt(4.0 * 386));
    EXPECT_FLOAT_EQ(0.1f + 0.2f, 0.3f);
    EXPECT_NEAR(1.000001, 1, 0.00001);
    EXPECT_STREQ(std::string("AoSoA").c_str(), "AoSoA");
    EXPECT_TRUE(true);
    ASSERT_EQ(size_t(5), size_t(2) + 3);
    EXPECT_EQ(-0.0f, 0.0f);
}
TEST(Shim, Fails)
{
    EXPECT_EQ(1, 2);
    EXPECT_FLOAT_EQ(1.0f, 1.00001f);
    EXPECT_FLOAT_EQ(-1e-30f, 1e-30f);
    ASSERT_TRUE(false);
    EXPECT_EQ(3, 4);  // not reached
}
int main()
{
    for (auto& t : testing_shim::registry()) {
        int b = testing_shim::failures();
        t.body();
        printf("%s.%s failures %d\n", t.suite, t.name, testing_shim::failures() - b);
    }
}
