// oracle/ref_gtests_main.cu -- runs THE REFERENCE'S OWN gtest files, included unmodified from the reference tree
// (tests/RXMesh_test/test_attribute.cu, test_for_each.cu, test_boundary.cu, test_export.cu, test_ev_diamond.cu), against the
// drop-in headers and librxmesh_b200.so.  googletest itself is third party and absent: ref_shim/user/gtest/gtest.h supplies
// TEST / EXPECT_* / ASSERT_*.  INPUT_DIR is `rxm_input/` (relative): the caller (tests/test_zz_reference_sources.py) runs
// the binary in a directory that holds rxm_input/<mesh>.obj written from the committed fixtures.  Test infrastructure only.
// Exit code = number of failed checks; prints gtest-style [ RUN ] / [ OK ] / [ FAILED ] lines.
#include "gtest/gtest.h"

#include "rxmesh/rxmesh_static.h"

#include "test_attribute.cu"
#include "test_boundary.cu"
#include "test_ev_diamond.cu"
#include "test_export.cu"
#include "test_for_each.cu"

int main(int argc, char** argv)
{
    rxmesh::rx_init(0);
    const std::string filter = argc > 1 ? argv[1] : "";
    int               ran = 0, failed_tests = 0;
    for (const auto& t : testing_shim::registry()) {
        const std::string full = std::string(t.suite) + "." + t.name;
        if (!filter.empty() && full.find(filter) == std::string::npos) continue;
        printf("[ RUN      ] %s\n", full.c_str());
        fflush(stdout);
        const int before = testing_shim::failures();
        t.body();
        ++ran;
        if (testing_shim::failures() != before) {
            ++failed_tests;
            printf("[  FAILED  ] %s\n", full.c_str());
        } else {
            printf("[       OK ] %s\n", full.c_str());
        }
        fflush(stdout);
    }
    printf("[==========] %d tests ran, %d failed\n", ran, failed_tests);
    return testing_shim::failures() > 255 ? 255 : testing_shim::failures();
}
