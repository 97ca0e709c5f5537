// TEST INFRASTRUCTURE ONLY. Minimal stand-in for boost::pool<> (Boost is not installed in this
// image) so the reference's own translation units compile unmodified into oracle/_ref/.
// boost::pool is an allocator with no arithmetic; malloc/free preserves every result.
#pragma once
#include <cstddef>
#include <cstdlib>
namespace boost {
struct default_user_allocator_new_delete {};
template <typename UserAllocator = default_user_allocator_new_delete>
class pool {
    std::size_t m_size;
public:
    explicit pool(std::size_t requested_size) : m_size(requested_size) {}
    void *malloc() { return std::malloc(m_size); }
    void free(void *p) { std::free(p); }
};
}
