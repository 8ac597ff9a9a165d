/* glibc>=2.26 defines CHAR_WIDTH, which the reference's vendored fmt
 * (src/thirdparty/spdlog/fmt/bundled/format.h:2198) uses as an identifier. */
#include <limits.h>
#include <climits>
#undef CHAR_WIDTH
