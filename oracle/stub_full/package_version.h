/* TEST INFRASTRUCTURE ONLY: the reference's build generates this header with cmake (src/version/CMakeLists.txt) */
#define PACKAGE_VERSION "0.7.4-oracle"
