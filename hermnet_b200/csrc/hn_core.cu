// ABI bookkeeping: version, last-error string, device query.
#include "hn_common.cuh"

namespace hn {
thread_local std::string g_last_error;
}

extern "C" int hn_abi_version(void) { return HN_ABI_VERSION; }
extern "C" const char *hn_last_error(void) { return hn::g_last_error.c_str(); }
extern "C" int hn_device_sm_count(void) { return hn::num_sms(); }
