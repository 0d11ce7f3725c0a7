#include "cp_oracle.h"
int cp_oracle_version(void){ return 1; }
