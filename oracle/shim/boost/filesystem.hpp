// NOT reference code. Stand-in for <boost/filesystem.hpp> (Boost is not installed):
// the reference only calls boost::filesystem::create_directories.
#pragma once
#include <filesystem>
namespace boost { namespace filesystem { using std::filesystem::create_directories; } }
