/* stand-in for netcdf.h: only the constants the reference headers use */
#ifndef GOMA_B200_ORACLE_NETCDF_STUB_H
#define GOMA_B200_ORACLE_NETCDF_STUB_H
#define NC_MAX_NAME 256
#define NC_MAX_DIMS 1024
#define NC_MAX_VAR_DIMS 1024
#define NC_NOERR 0
#define NC_NOWRITE 0
#define NC_WRITE 1
#define NC_CLOBBER 0
#define NC_GLOBAL (-1)
#define NC_SHARE 0x0800
typedef int nc_type;
#define NC_INT 4
#define NC_DOUBLE 6
#define NC_CHAR 2
#endif
