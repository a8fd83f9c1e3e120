// TEST INFRASTRUCTURE ONLY (oracle/). The reference includes this header but
// calls nothing from it.
#ifndef KMSC_ORACLE_SHIM_STR_JOIN_H_
#define KMSC_ORACLE_SHIM_STR_JOIN_H_
#endif
