/* placeholder translation unit for the config-5 (KmerCountExact) oracle; filled in when that row is built */
typedef int kcount_oracle_placeholder_t;
