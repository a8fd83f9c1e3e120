// TEST INFRASTRUCTURE ONLY (oracle/). boost::asio::post(pool, f).
#ifndef KMSC_ORACLE_SHIM_POST_HPP_
#define KMSC_ORACLE_SHIM_POST_HPP_
#include <utility>
#include "boost/asio/thread_pool.hpp"
namespace boost { namespace asio {
template <typename F>
void post(thread_pool& pool, F&& f) { pool.Post(std::forward<F>(f)); }
}}
#endif
