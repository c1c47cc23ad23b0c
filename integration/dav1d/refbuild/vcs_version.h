/* stand-in for the meson-generated vcs_version.h */
#define DAV1D_VERSION "1.5.4-oracle-ref"
