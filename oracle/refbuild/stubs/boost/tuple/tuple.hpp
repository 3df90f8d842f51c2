// Stand-in for boost/tuple/tuple.hpp (Boost is absent): the member get<N>() and make_tuple the reference uses.
#pragma once
#include <tuple>
namespace boost {
template <class... T>
struct tuple : std::tuple<T...> {
    using std::tuple<T...>::tuple;
    template <int N> typename std::tuple_element<N, std::tuple<T...>>::type& get() { return std::get<N>(*this); }
    template <int N> const typename std::tuple_element<N, std::tuple<T...>>::type& get() const { return std::get<N>(*this); }
};
template <class... T> tuple<T...> make_tuple(T... v) { return tuple<T...>(v...); }
}
