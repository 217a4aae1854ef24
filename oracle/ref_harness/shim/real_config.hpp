/* Hand-written stand-in for the header the reference's configure script would
 * generate from real_config.hpp.in.  TEST INFRASTRUCTURE ONLY: it exists so that
 * oracle/Makefile can compile the reference's own sources, in place under
 * /root/reference/src, into oracle/_ref/ without running the reference's build
 * system.  Values follow configure.in:247-272 (reader types, EPSILON,
 * MATCHING_SINGLE_INSTANTIATION, _FILE_OFFSET_BITS) for a Linux/x86_64 host. */
#ifndef ORACLE_SHIM_REAL_CONFIG_HPP
#define ORACLE_SHIM_REAL_CONFIG_HPP

#define EPSILON 1e-6
#define SLOW_UNIQUE_FASTA_READER_TYPE FastAReader
#define SLOW_UNIQUE_FASTQ_READER_TYPE FastQReader
#define FAST_UNIQUE_FASTA_READER_TYPE FastFileDecoder
#define FAST_UNIQUE_FASTQ_READER_TYPE FastQualityFileDecoder
#define SLOW_ALL_FASTA_READER_TYPE FastAReader
#define SLOW_ALL_FASTQ_READER_TYPE FastQReader
#define MATCHING_SINGLE_INSTANTIATION 1

#define HAVE_AIO_H 1
#define HAVE_DIRENT_H 1
#define HAVE_PTHREADS 1
#define HAVE_SEM_DESTROY 1
#define HAVE_SEM_INIT 1
#define HAVE_SEM_POST 1
#define HAVE_SEM_WAIT 1
#define HAVE_STDINT_H 1
#define HAVE_STDLIB_H 1
#define HAVE_STRING_H 1
#define HAVE_SYS_STAT_H 1
#define HAVE_SYS_TIME_H 1
#define HAVE_SYS_TYPES_H 1
#define HAVE_UNISTD_H 1
#define HAVE_LINUX_SYSCTL_H 1
#define HAVE_x86_64 1

#define PACKAGE "real"
#define PACKAGE_NAME "real"
#define PACKAGE_VERSION "0.0.31"
#define PACKAGE_STRING "real 0.0.31"
#define VERSION "0.0.31"
#ifndef _FILE_OFFSET_BITS
#define _FILE_OFFSET_BITS 64
#endif

#endif
